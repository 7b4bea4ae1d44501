#!/bin/bash
# round 2, GPU call A: diagnose graph capture at HEAD, then the round-1 "first call" script
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
nproc
timeout 300 python scripts/diag_capture.py > gpurun_out/diag_capture.log 2>&1; echo "diag rc=$?"; tail -40 gpurun_out/diag_capture.log | cut -c1-400
# (then ran the round-1 hand-over script, since deleted: experimental-kernel tests, suite, bench A/B)
