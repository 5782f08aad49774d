#!/bin/bash
# tests of the new matcher paths + everything, the default bench, the ncu capture of the tensor-core matcher
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_match_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/r2_run3_match.log 2>&1; echo "match tests exit: $?"; tail -15 gpurun_out/r2_run3_match.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_match_gpu.py > gpurun_out/r2_run3_all.log 2>&1; echo "pytest exit: $?"; tail -5 gpurun_out/r2_run3_all.log
( time timeout 1500 python bench.py > gpurun_out/r2_bench_v1.json 2> gpurun_out/r2_bench_v1.err ) 2> gpurun_out/r2_bench_v1.time; echo "bench exit: $?"; tail -3 gpurun_out/r2_bench_v1.err; cat gpurun_out/r2_bench_v1.time
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_v1.json').read().strip().splitlines()[-1])
    print('BA', d['value'], d['value_run'], d['ba_ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'])
    m=d['match']; print('MATCH', m['value'], m['e2e']['value'], m['roofline']['frac'])
    for k,v in d.get('extras',{}).items(): print(k, json.dumps(v)[:600])
    print(d.get('cpu_baseline'), m.get('cpu_baseline'))
except Exception as e: print('parse failed', e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^bf_top2_tc' -c 1 -f -o gpurun_out/r2_full_tc python scripts/prof_match.py > gpurun_out/r2_ncu_full_tc.log 2>&1; echo "ncu full tc: $?"
