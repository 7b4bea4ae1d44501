#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layers_gpu.py -q -m gpu > gpurun_out/layers_test.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
grep -E "^(FAILED|ERROR)|passed|failed|Max abs|Mismatched|err_msg|AssertionError: .+" gpurun_out/layers_test.log | head -60
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -c 3000 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
