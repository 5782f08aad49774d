#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_match_gpu.py -m gpu -q --timeout 300 -x > gpurun_out/r2_run23_match.log 2>&1; echo "match pytest exit: $?"; tail -15 gpurun_out/r2_run23_match.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_run23_all.log 2>&1; echo "pytest exit: $?"; tail -12 gpurun_out/r2_run23_all.log
( time timeout 1500 python bench.py > gpurun_out/r2_bench_v9.json 2> gpurun_out/r2_bench_v9.err ) 2> gpurun_out/r2_bench_v9.time; echo "bench exit: $?"; tail -3 gpurun_out/r2_bench_v9.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_v9.json').read().strip().splitlines()[-1])
    print('BA', d['value'], d['value_run'], d['ba_ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'])
    m=d['match']; print('MATCH', m['value'], m['e2e']['value'], m['roofline']['frac'])
    g=d['extras']['match_guided']; print('guided', g['value'], g['device_ms'], g['distance_kernel_ms'], g['e2e']['value'])
    for k in ('match_hamming_akaze61','match_hamming_orb32','match_float'):
        e=d['extras'][k]; print(k, e['kernel'], e['value'], e['distance_kernel_ms'], e['roofline']['frac'])
except Exception as e: print('parse failed', e)
PY
