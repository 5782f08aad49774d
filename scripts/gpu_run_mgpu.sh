#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/mgpu_gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x --timeout 800 > gpurun_out/pytest_mgpu.log 2>&1; echo "exit: $?" >> gpurun_out/pytest_mgpu.log
tail -15 gpurun_out/pytest_mgpu.log
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "exit bench n=$N: $?"
tail -c 2500 gpurun_out/bench_c4_n$N.json; tail -5 gpurun_out/bench_c4_n$N.err
