#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; tail -2 gpurun_out/bench_n2.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
    print('N=2', d['value'], d['ms_per_step'], d['n_gpus'], d['config'], d['e2e']['value'])
    b=d['breakdown_ms_per_step']
    for k,v in list(b.items())[:4]: print('   %-32s %.3f (%d)'%(k,v['ms_per_step'],v['calls_per_step']))
except Exception as e: print('N=2 parse error', e)
PY
