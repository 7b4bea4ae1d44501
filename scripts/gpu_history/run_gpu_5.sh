#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_smoke.py > gpurun_out/debug_smoke.log 2>&1; tail -80 gpurun_out/debug_smoke.log
timeout 400 python -m pytest tests/test_gemm_gpu.py -q -m gpu > gpurun_out/gemm_test.log 2>&1
echo "gemm test rc=$?"; grep -E "passed|failed|^FAILED|Max abs|Mismatch" gpurun_out/gemm_test.log | head -40
