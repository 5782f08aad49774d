#!/bin/bash
mkdir -p gpurun_out
for nb in 3 4 5; do
OSFM_BA_LIN_NB=$nb timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']; print('NB $nb: BA', round(d['ba_ms_per_step'],3), 'ms  linearize', round(k['ba_linearize']['ms'],4), 'ms/launch')"
done
for nb in 4 5; do
OSFM_BA_LIN_NB=$nb timeout 600 python -m pytest tests/test_ba_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -1
done
