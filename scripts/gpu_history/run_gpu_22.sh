#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_graph2.py 16 > gpurun_out/debug_graph2.log 2>&1; grep -v "^Search\|^CUDA kernel\|^For debugging\|^Compile\|^$" gpurun_out/debug_graph2.log | tail -16 | cut -c1-300
