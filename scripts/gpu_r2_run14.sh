#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ba_gpu.py tests/test_ba_parity_scale.py tests/test_bundle_reference.py tests/test_reconstruction_bundle.py tests/test_reconstruction_alignment.py -m gpu -q --timeout 600 > gpurun_out/r2_run14_ba.log 2>&1; echo "ba pytest exit: $?"; tail -8 gpurun_out/r2_run14_ba.log
python - <<'PY'
import sys; sys.path.insert(0,'.')
import bench
from opensfm_b200 import bundle
pb, feats, pairs, w = bench.build_workload("c4")
bundle.solve(pb)
for i in range(3):
    r = bundle.solve(pb); s = r["summary"]
    print("device %.2f ms run %.2f schur %.2f pcg %.2f lin %.2f backsub %.2f" % (s["time_device_ms"], 1e3*s["time_run_s"], s["time_schur_ms"], s["time_pcg_ms"], s["time_linearize_ms"], s["time_backsub_ms"]))
PY
KF='regex:^(ba_|pcg_|sp_|side_)'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" --csv --log-file gpurun_out/r2_launches_ba_v6.csv python scripts/trace_ba.py c4 > /dev/null 2>&1; echo "ncu list: $?"
python scripts/summarize_launches.py gpurun_out/r2_launches_ba_v6.csv | head -24
