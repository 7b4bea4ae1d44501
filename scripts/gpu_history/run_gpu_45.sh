#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
STRESS_ITERS=600 timeout 300 python scripts/stress_tc2.py 2>&1 | tail -3 | cut -c1-150
timeout 300 python scripts/bench_gemm.py 2>&1 | tail -17
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench45.err | tail -1 > gpurun_out/bench45.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench45.json'))
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])
PY
