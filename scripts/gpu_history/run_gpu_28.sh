#!/bin/bash
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -4
timeout 300 python scripts/debug_tc_trace.py 2>&1 | grep -v "chunks\|tiles\|epi tmem\|epi lds\|epi stg\|epi stats" | head -60
timeout 300 python scripts/bench_gemm.py 2>&1 | tail -17
