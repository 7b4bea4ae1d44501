#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -8
timeout 300 python scripts/debug_tc_trace.py 2>&1 | head -56
timeout 300 python scripts/bench_gemm.py 2>&1 | tail -40
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench26.err | tail -1 > gpurun_out/bench26.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench26.json'))
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])
PY
grep "ms (" gpurun_out/bench26.err | head -14
