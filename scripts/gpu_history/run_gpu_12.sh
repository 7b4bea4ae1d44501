#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; cat gpurun_out/bench_gemm.log
