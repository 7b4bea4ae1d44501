#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_tc_trace.py > gpurun_out/tc_trace.log 2>&1; cat gpurun_out/tc_trace.log
