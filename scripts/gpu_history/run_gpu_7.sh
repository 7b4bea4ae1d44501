#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_smoke2.py > gpurun_out/debug_smoke2.log 2>&1; tail -30 gpurun_out/debug_smoke2.log
