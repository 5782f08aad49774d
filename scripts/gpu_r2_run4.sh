#!/bin/bash
# all GPU tests, the default bench, A/B of the bulk-copy colnorm kernel, ncu launch list + full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_run4_all.log 2>&1; echo "pytest exit: $?"; tail -8 gpurun_out/r2_run4_all.log
( time timeout 1500 python bench.py > gpurun_out/r2_bench_v2.json 2> gpurun_out/r2_bench_v2.err ) 2> gpurun_out/r2_bench_v2.time; echo "bench exit: $?"; tail -3 gpurun_out/r2_bench_v2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_v2.json').read().strip().splitlines()[-1])
    print('BA', d['value'], d['value_run'], d['ba_ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'])
    print({k:(v.get('ms'), v.get('share')) for k,v in d['roofline']['kernels'].items()})
    m=d['match']; print('MATCH', m['value'], m['e2e']['value'], m['roofline']['frac'])
    g=d['extras']['match_guided']; print('guided', g['value'], g['device_ms'], g['distance_kernel_ms'], g['e2e']['value'])
except Exception as e: print('parse failed', e)
PY
OSFM_BA_COLNORM_TMA=0 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v2_notma.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_v2_notma.json').read().strip().splitlines()[-1]); print('no-tma BA', d['value'], d['ba_ms_per_step'])"
KF='regex:^(ba_|bf_|pcg_|tc_|ord_|bsr_|pad_rows|side_|epi_|widen)'
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" --csv --log-file gpurun_out/r2_launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_under_ncu.log 2>&1; echo "ncu list: $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:^(ba_point_blocks|ba_schur_mma|ba_linearize|pcg_pipelined|ba_colnorm_grad_tma|ba_colnorm_grad_points|ba_backsub_rows|ba_model_change)' -c 10 -f -o gpurun_out/r2_full_ba python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_full_ba.log 2>&1; echo "ncu full ba: $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:^bf_top2_tc' -c 1 -f -o gpurun_out/r2_full_tc python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_full_tc.log 2>&1; echo "ncu full tc: $?"
