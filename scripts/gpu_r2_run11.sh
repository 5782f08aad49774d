#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/check_pipe.py > gpurun_out/r2_check_pipe.log 2>&1; rc=$?; echo "check_pipe exit: $rc"; tail -8 gpurun_out/r2_check_pipe.log
if [ $rc -ne 0 ]; then
  timeout 600 compute-sanitizer --tool memcheck python -c "
import sys; sys.path.insert(0, '/root/repo')
from opensfm_b200 import bundle, synthetic as syn
sc = syn.cube_scene(30, 4000, 1.0, with_descriptors=False, max_obs_per_point=8)
r = bundle.solve(syn.scene_to_problem(sc)); print(r['summary']['final_cost'])
" > gpurun_out/r2_sanitizer.log 2>&1; head -60 gpurun_out/r2_sanitizer.log
  exit 1
fi
timeout 300 python scripts/trace_ba.py c4 > gpurun_out/r2_trace_c4_pipe.log 2>&1; grep -A2 "ba_schur_pipe" gpurun_out/r2_trace_c4_pipe.log | head -4; tail -1 gpurun_out/r2_trace_c4_pipe.log
python - <<'PY'
import sys; sys.path.insert(0,'.')
import bench
from opensfm_b200 import bundle
pb, feats, pairs, w = bench.build_workload("c4")
bundle.solve(pb)
for i in range(3):
    r = bundle.solve(pb); s = r["summary"]
    print("no-trace: device %.2f ms schur %.2f pcg %.2f lin %.2f" % (s["time_device_ms"], s["time_schur_ms"], s["time_pcg_ms"], s["time_linearize_ms"]))
PY
timeout 1200 python -m pytest tests/test_ba_gpu.py tests/test_ba_parity_scale.py tests/test_bundle_reference.py tests/test_reconstruction_bundle.py tests/test_reconstruction_alignment.py -m gpu -q --timeout 600 > gpurun_out/r2_run11_ba.log 2>&1; echo "ba pytest exit: $?"; tail -8 gpurun_out/r2_run11_ba.log
