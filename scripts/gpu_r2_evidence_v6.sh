#!/bin/bash
# round-2 evidence: launch list + full ncu captures of the dominant kernels, per-phase trace, reference arm
mkdir -p gpurun_out
KF='regex:^(ba_|bf_|pcg_|tc_|h8_|sp_|ord_|bsr_|pad_rows|side_|epi_|widen)'
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" --csv --log-file gpurun_out/r2_launches_c4_v6.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_under_ncu.log 2>&1; echo "ncu list: $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:^(ba_point_blocks|ba_schur_pipe|ba_linearize|pcg_pipelined|ba_colnorm_grad_chunks|ba_colnorm_grad_tma|ba_backsub_rows|ba_model_change)' -c 9 -f -o gpurun_out/r2_full_ba_v6 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_full_ba.log 2>&1; echo "ncu full ba: $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:^bf_top2_tc' -c 1 -f -o gpurun_out/r2_full_tc_v6 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_full_tc.log 2>&1; echo "ncu full tc: $?"
timeout 300 python scripts/trace_ba.py c4 > gpurun_out/r2_trace_c4_v6.log 2>&1; echo "trace: $?"
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_v6.json 2> gpurun_out/r2_bench_reference_v6.err ) 2> gpurun_out/r2_bench_reference_v6.time; echo "reference arm: $?"; tail -c 600 gpurun_out/r2_bench_reference_v6.json
( time timeout 1500 python bench.py > gpurun_out/r2_bench_v7.json 2> gpurun_out/r2_bench_v7.err ) 2> gpurun_out/r2_bench_v7.time; echo "bench exit: $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_v7.json').read().strip().splitlines()[-1])
print('BA', d['value'], d['value_run'], d['ba_ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'])
print({k:(round(v.get('ms',0),3), round(v.get('share',0),3)) for k,v in d['roofline']['kernels'].items()})
PY
