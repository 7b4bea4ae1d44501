#!/bin/bash
# round 2, GPU call U: geometry-ahead parity test at the learning-rate floor, full suite
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests/test_train_step_gpu.py -m gpu -x -q -s > gpurun_out/u_train.log 2>&1; echo "train tests rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/u_train.log)"; grep -E "^FAILED|^ERROR|^losses|^E  " gpurun_out/u_train.log | cut -c1-500
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/u_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/u_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/u_suite.log | cut -c1-300
