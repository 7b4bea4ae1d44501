#!/bin/bash
# round 2, GPU call D: suite with the new kNN / selection / pooling / predictor / box-sampling tests, role-cycle trace
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/d_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/d_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/d_suite.log | cut -c1-400
PN2_LIB=$PWD/open3d-pointnet2-semantic3d_b200/lib/libpn2_b200_trace.so timeout 120 python scripts/debug_tc_trace.py > gpurun_out/d_trace.log 2>&1; echo "trace rc=$? t=$((SECONDS-T0))"; cat gpurun_out/d_trace.log | cut -c1-200
