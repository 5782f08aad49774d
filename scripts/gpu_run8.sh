#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit: $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for w in c2 c4; do
timeout 1500 python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "exit $w: $?"
python - <<PY
import json
f="gpurun_out/bench_$w.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "BA", d["value"], d["ba_ms_per_step"], d["roofline"]["kernels"], "e2e", d["e2e"]["value"], "MATCH", d["match"]["value"], d["match"]["roofline"]["frac"], d["match"]["e2e"]["value"])
except Exception as e: print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-2000:])
PY
done
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:ba_schur_seg|pcg_persistent" -c 2 -f -o gpurun_out/prof_ba python scripts/prof_ba.py c4 > gpurun_out/ncu_ba.log 2>&1; echo "ncu ba: $?"
