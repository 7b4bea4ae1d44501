#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_graph2.py > gpurun_out/debug_graph2.log 2>&1; tail -40 gpurun_out/debug_graph2.log | cut -c1-250
