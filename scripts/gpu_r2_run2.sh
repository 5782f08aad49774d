#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_bundle_reference.py tests/test_reconstruction_bundle.py -m gpu -q -x --timeout 600 > gpurun_out/r2_run2_new.log 2>&1; echo "new tests exit: $?"; tail -25 gpurun_out/r2_run2_new.log
timeout 1500 python -m pytest tests/test_bundle_reference.py tests/test_reconstruction_bundle.py -m gpu -q --timeout 600 > gpurun_out/r2_run2_new_all.log 2>&1; tail -15 gpurun_out/r2_run2_new_all.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_ba_parity_scale.py --deselect tests/test_bundle_reference.py --deselect tests/test_reconstruction_bundle.py > gpurun_out/r2_run2_all.log 2>&1; echo "pytest exit: $?"; tail -3 gpurun_out/r2_run2_all.log
