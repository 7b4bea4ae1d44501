#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/debug_wgrad.py > gpurun_out/debug_wgrad.log 2>&1; tail -30 gpurun_out/debug_wgrad.log
