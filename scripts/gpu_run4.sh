#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit: $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python scripts/prof_match.py > gpurun_out/prof_match.log 2>&1; tail -2 gpurun_out/prof_match.log
timeout 900 python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "exit c2: $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_c2.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "BA", d["value"], d["ba_ms_per_step"], d["roofline"]["kernels"], "MATCH", d["match"]["value"], d["match"]["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
timeout 1500 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "exit c4: $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_c4.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "BA", d["value"], d["ba_ms_per_step"], d["roofline"]["kernels"], "e2e", d["e2e"]["value"], "MATCH", d["match"]["value"], d["match"]["roofline"]["frac"], d["match"]["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
# ncu: launch list of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches: $?"
# ncu: full captures of the top kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bf_top2_tc -s 1 -c 1 -f -o gpurun_out/prof_tc python scripts/prof_match.py > gpurun_out/ncu_tc.log 2>&1; echo "ncu tc: $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:ba_schur|pcg_persistent|ba_linearize" -c 4 -f -o gpurun_out/prof_ba python scripts/prof_ba.py c4 > gpurun_out/ncu_ba.log 2>&1; echo "ncu ba: $?"
ls -la gpurun_out | head -30
