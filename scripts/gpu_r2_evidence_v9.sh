#!/bin/bash
# round-2 final evidence: launch list + full ncu captures of the BA kernels, per-phase trace (state of bench v9)
mkdir -p gpurun_out
KF='regex:^(ba_|bf_|pcg_|tc_|h8_|sp_|ord_|bsr_|pad_rows|side_|epi_|widen)'
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" --csv --log-file gpurun_out/r2_launches_c4_v9.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_under_ncu.log 2>&1; echo "ncu list: $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:^(ba_point_blocks|ba_schur_pipe|ba_linearize|pcg_pipelined|ba_colnorm_grad_chunks|ba_backsub_rows|ba_model_change_alg)' -s 8 -c 8 -f -o gpurun_out/r2_full_ba_v9 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_full_ba.log 2>&1; echo "ncu full ba: $?"
timeout 300 python scripts/trace_ba.py c4 > gpurun_out/r2_trace_c4_v9.log 2>&1; echo "trace: $?"
