#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/trace_ba.py c4 > gpurun_out/r2_trace_c4_pipe.log 2>&1; grep -A2 "ba_schur_pipe" gpurun_out/r2_trace_c4_pipe.log | head -8; tail -1 gpurun_out/r2_trace_c4_pipe.log
timeout 1200 python -m pytest tests/test_ba_gpu.py tests/test_ba_parity_scale.py tests/test_bundle_reference.py tests/test_reconstruction_bundle.py tests/test_reconstruction_alignment.py -m gpu -q --timeout 600 > gpurun_out/r2_run10_ba.log 2>&1; echo "ba pytest exit: $?"; tail -15 gpurun_out/r2_run10_ba.log
