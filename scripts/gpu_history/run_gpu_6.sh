#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_chain.py > gpurun_out/debug_chain.log 2>&1; tail -40 gpurun_out/debug_chain.log
timeout 400 python -m pytest tests/test_gemm_gpu.py -q -m gpu > gpurun_out/gemm_test.log 2>&1
echo "gemm test rc=$?"; grep -E "passed|failed|^FAILED|Max abs|Mismatch" gpurun_out/gemm_test.log | head -40
timeout 400 python -m pytest tests/test_ops_gpu.py -q -m gpu > gpurun_out/ops_test.log 2>&1
echo "ops test rc=$?"; grep -E "passed|failed|^FAILED|Max abs|Mismatch" gpurun_out/ops_test.log | head -20
