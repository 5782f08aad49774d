#!/bin/bash
# Round evidence on one B200: tests, the default bench (+ cpu_baseline), the reference arm, the ncu launch
# list of the bench command and ncu --set full captures of its dominant kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit: $?"
tail -3 gpurun_out/pytest_gpu.log
( time timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time; echo "bench exit: $?"
( time timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2> gpurun_out/bench_reference.time; echo "reference exit: $?"
KF='regex:^(ba_|bf_|pcg_|tc_|ord_|bsr_|pad_rows)'
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list: $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:^(ba_point_blocks|ba_schur_mma|ba_linearize|pcg_pipelined|ba_colnorm_grad_seg|ba_colnorm_grad_points|ba_backsub_rows)$' -c 10 -f -o gpurun_out/full_ba python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_ba.log 2>&1; echo "ncu full ba: $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:^bf_top2_tc$' -c 1 -f -o gpurun_out/full_tc python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_tc.log 2>&1; echo "ncu full tc: $?"
cat gpurun_out/bench_default.time gpurun_out/bench_reference.time
tail -c 1500 gpurun_out/bench_reference.json
