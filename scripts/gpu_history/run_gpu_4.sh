#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_model.py > gpurun_out/debug_model.log 2>&1
cat gpurun_out/debug_model.log | tail -70
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x > gpurun_out/gemm_test.log 2>&1
echo "gemm test rc=$?"; grep -E "passed|failed|Error|Max abs|Mismatch|^FAILED|^E  " gpurun_out/gemm_test.log | head -30
