#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -12
timeout 300 python scripts/debug_tc_trace.py 2>&1 | grep -A8 "wgrad M" | head -60
timeout 300 python scripts/bench_gemm.py 2>&1 | tail -17
