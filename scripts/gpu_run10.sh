#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit: $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for seg in 0 1; do
OSFM_BA_SEGMENT_SCHUR=$seg timeout 1500 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_seg$seg.json 2> gpurun_out/bench_c4_seg$seg.err; echo "exit seg$seg: $?"
python - <<PY
import json
f="gpurun_out/bench_c4_seg$seg.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "BA", d["value"], d["ba_ms_per_step"], d["roofline"]["kernels"], "e2e", d["e2e"]["value"])
except Exception as e: print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-2000:])
PY
done
OSFM_BA_SEGMENT_SCHUR=1 timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:ba_schur_seg|ba_obs_rows|ba_point_blocks" -c 3 -f -o gpurun_out/prof_seg python scripts/prof_ba.py c4 > gpurun_out/ncu_seg.log 2>&1; echo "ncu seg: $?"
