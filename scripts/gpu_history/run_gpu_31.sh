#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -6
PN2_TC_TMA=0 timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -3
timeout 300 python scripts/debug_tc_trace.py 2>&1 | grep -v "chunks\|tiles\|epi tmem\|epi lds\|epi stg\|epi stats" | head -60
timeout 300 python scripts/bench_gemm.py 2>&1 | tail -17
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench31.err | tail -1 > gpurun_out/bench31.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench31.json'))
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])
PY
grep "ms (" gpurun_out/bench31.err | head -14
