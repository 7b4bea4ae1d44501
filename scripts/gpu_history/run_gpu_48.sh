#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench48.err | tail -1 > gpurun_out/bench48.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench48.json'))
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])
b=d['breakdown_ms_per_step']
for k,v in list(b.items())[:14]: print('   %-32s %.3f (%d)'%(k,v['ms_per_step'],v['calls_per_step']))
PY
