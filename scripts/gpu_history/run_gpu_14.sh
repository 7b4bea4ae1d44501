#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gemm_gpu.py -q -m gpu > gpurun_out/gemm_test.log 2>&1
echo "gemm test rc=$?"; grep -E "passed|failed|^FAILED|Max abs|Mismatch" gpurun_out/gemm_test.log | head -30
timeout 600 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; cat gpurun_out/bench_gemm.log
timeout 300 python scripts/debug_tc_trace.py > gpurun_out/tc_trace.log 2>&1; head -14 gpurun_out/tc_trace.log
