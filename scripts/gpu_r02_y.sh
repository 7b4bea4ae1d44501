#!/bin/bash
# round 2, GPU call Y: compute-sanitizer over the final build (programmatic dependent launch, float4 paths, skinny wgrad,
# the Trainer with its side streams)
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_ops.py > gpurun_out/y_memcheck.log 2>&1; echo "memcheck rc=$? t=$((SECONDS-T0))"; tail -4 gpurun_out/y_memcheck.log | cut -c1-200
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_ops.py gemm layers > gpurun_out/y_racecheck.log 2>&1; echo "racecheck(gemm, layers) rc=$? t=$((SECONDS-T0))"; tail -4 gpurun_out/y_racecheck.log | cut -c1-200
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 python scripts/sanitize_ops.py gemm layers trainer > gpurun_out/y_synccheck.log 2>&1; echo "synccheck rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/y_synccheck.log | cut -c1-200
