#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/check_pipe.py > gpurun_out/r2_check_pipe.log 2>&1; rc=$?; echo "check_pipe exit: $rc"; tail -8 gpurun_out/r2_check_pipe.log
if [ $rc -ne 0 ]; then
  timeout 600 compute-sanitizer --tool memcheck python -c "
import sys; sys.path.insert(0, '/root/repo')
from opensfm_b200 import bundle, synthetic as syn
sc = syn.cube_scene(30, 4000, 1.0, with_descriptors=False, max_obs_per_point=8)
r = bundle.solve(syn.scene_to_problem(sc)); print(r['summary']['final_cost'])
" > gpurun_out/r2_sanitizer.log 2>&1; head -60 gpurun_out/r2_sanitizer.log
  exit 1
fi
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_ba_parity_scale.py -m gpu -q --timeout 600 -x > gpurun_out/r2_run9_ba.log 2>&1; echo "ba pytest exit: $?"; tail -15 gpurun_out/r2_run9_ba.log
timeout 300 python scripts/trace_ba.py c4 > gpurun_out/r2_trace_c4_pipe.log 2>&1; tail -4 gpurun_out/r2_trace_c4_pipe.log
OSFM_BA_SCHUR_PIPE=0 timeout 300 python scripts/trace_ba.py c4 2>&1 | tail -1
