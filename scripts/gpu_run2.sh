#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit: $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --workload c2 --steps 2 --warmup 1 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "exit c2: $?"
tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
timeout 1500 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "exit c4: $?"
tail -c 4000 gpurun_out/bench_c4.json; tail -5 gpurun_out/bench_c4.err
