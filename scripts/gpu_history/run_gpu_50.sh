#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench50.err | tail -1 > gpurun_out/bench50.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench50.json'))
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])
b=d['breakdown_ms_per_step']
for k,v in list(b.items())[:8]: print('   %-32s %.3f (%d)'%(k,v['ms_per_step'],v['calls_per_step']))
PY
PN2_FPS_PRUNE=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no-prune', d['ms_per_step'], d['breakdown_ms_per_step']['pn2_fps'])"
