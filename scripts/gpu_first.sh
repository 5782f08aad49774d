#!/bin/bash
# bring-up on the GPU box: isolated stages with timeouts, logs into gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
for s in simt ba tc; do
  timeout 300 python scripts/gpu_probe1.py $s > gpurun_out/probe_$s.log 2>&1; echo "exit $s: $?" >> gpurun_out/probe_$s.log
done
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit: $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/probe_*.log gpurun_out/pytest_gpu.log
