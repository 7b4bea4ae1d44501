#!/bin/bash
timeout 900 python -m pytest tests/test_gemm_gpu.py -q -k "padded_rows_many or dgrad" 2>&1 | tail -25
echo "---- PN2_TC_TMA=1 (A loads only)"
PN2_TC_TMA=1 timeout 900 python -m pytest tests/test_gemm_gpu.py -q -k "padded_rows_many" 2>&1 | tail -8
echo "---- PN2_TC_TMA=2 (Y stores only)"
PN2_TC_TMA=2 timeout 900 python -m pytest tests/test_gemm_gpu.py -q -k "padded_rows_many" 2>&1 | tail -8
