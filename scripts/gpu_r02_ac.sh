#!/bin/bash
# round 2, GPU call AC: the new full-size geometry-ahead equivalence test + the train-step file, sanitizers on the final build
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests/test_train_step_gpu.py -m gpu -x -q -s > gpurun_out/ac_train.log 2>&1; echo "train tests rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/ac_train.log)"; grep -E "^FAILED|^ERROR|^E  " gpurun_out/ac_train.log | cut -c1-400
bash scripts/gpu_r02_y.sh
