#!/bin/bash
# multi-GPU weak scaling check (2 ranks) + reference arm + default single-GPU line
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; tail -3 gpurun_out/bench_n2.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
    print('N=2', d['value'], d['ms_per_step'], d['n_gpus'], d['config'], d['e2e']['value'])
except Exception as e: print('N=2 parse error', e)
PY
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"; tail -1 gpurun_out/bench_ref_n2.json | cut -c1-400; tail -4 gpurun_out/bench_ref_n2.err
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default rc=$?"; tail -4 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('N=1', d['value'], d['ms_per_step'], d['e2e'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])
PY
