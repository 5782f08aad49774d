#!/bin/bash
mkdir -p gpurun_out
nproc; nvidia-smi -L | head -2
timeout 1500 python -m pytest tests/test_ba_parity_scale.py tests/test_match_gpu.py tests/test_golden.py -m gpu -q -s --timeout 900 > gpurun_out/r2_run1_new.log 2>&1; echo "new tests exit: $?"; grep -E "oracle:|max \|delta|passed|failed|Error|error" gpurun_out/r2_run1_new.log | tail -30
timeout 600 python scripts/prof_simt.py 2>&1 | tail -6
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_ba_parity_scale.py > gpurun_out/r2_run1_all.log 2>&1; echo "pytest exit: $?"; tail -3 gpurun_out/r2_run1_all.log
