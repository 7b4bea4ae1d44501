#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python scripts/bench_gemm.py 2>&1 | grep "131, 128\|M,K,N"
PN2_WGRAD_TAIL=1 timeout 300 python scripts/bench_gemm.py 2>&1 | grep "131, 128"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench49.err | tail -1 > gpurun_out/bench49.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench49.json'))
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])
PY
