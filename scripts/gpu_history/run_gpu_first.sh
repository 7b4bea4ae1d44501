#!/bin/bash
# first GPU contact: op parity tests
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/ops_test.log
