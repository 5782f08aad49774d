#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -2
for b in 1 0; do
OSFM_BA_PCG_B128=$b timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']; print('B128 $b: BA', round(d['ba_ms_per_step'],3), 'ms  pcg', round(k['pcg']['ms'],3), 'ms', k['pcg']['iterations'], 'its; e2e', d['e2e']['value'])"
done
