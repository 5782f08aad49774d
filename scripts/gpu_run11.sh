#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit: $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
OSFM_BA_TRACE=1 timeout 600 python scripts/prof_host.py c4 2>&1 | tail -24
timeout 1500 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "exit bench: $?"
python - <<PY
import json
f="gpurun_out/bench_c4.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "BA", d["value"], d["ba_ms_per_step"], d["roofline"]["kernels"], "e2e", d["e2e"])
except Exception as e: print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-2000:])
PY
