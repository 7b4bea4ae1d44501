#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_graph2.py 16 > gpurun_out/debug_graph2.log 2>&1; tail -8 gpurun_out/debug_graph2.log | cut -c1-1500
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu > gpurun_out/gemm_test.log 2>&1
echo "gemm test rc=$?"; grep -E "passed|failed|^FAILED|Max abs|Mismatch" gpurun_out/gemm_test.log | head
timeout 600 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; cut -c1-110 gpurun_out/bench_gemm.log
