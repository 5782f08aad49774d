#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit: $?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit: $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('BA', d['value'], d['ba_ms_per_step'], 'e2e', d['e2e']['value'], 'traffic', d['roofline']['traffic'], d['roofline']['frac']); m=d['match']; print('MATCH', m['value'], m['e2e']['value'], m['roofline']['frac'], m['roofline']['traffic']); print(d['cpu_baseline']['value'], m['cpu_baseline']['value'])"
