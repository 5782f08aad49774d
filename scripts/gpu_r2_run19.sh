#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}_v9.json 2> gpurun_out/r2_bench_n${N}_v9.err ) 2> gpurun_out/r2_bench_n${N}_v9.time; echo "bench N=$N exit: $?"; tail -3 gpurun_out/r2_bench_n${N}_v9.err | cut -c1-300
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2_bench_n%s_v9.json'%n).read().strip().splitlines()[-1])
    print(n,'BA', d['value'], d['value_run'], d['ba_ms_per_step'], d['ba_run_ms_per_step'], 'e2e', d['e2e']['value'])
    print('  kern', {k:(round(v['ms'],3), round(v['share'],3)) for k,v in d['roofline']['kernels'].items()})
    m=d['match']; print('  MATCH', m['value'], m['e2e']['value'], m['roofline']['frac'], m.get('images_resident_on_rank0'))
except Exception as e: print('parse failed', e)
PY
