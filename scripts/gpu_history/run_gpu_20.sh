#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_graph2.py 16 > gpurun_out/debug_graph2.log 2>&1; tail -12 gpurun_out/debug_graph2.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; tail -3 gpurun_out/bench_graph.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_graph.json"))
print("value %.4g ms/step %.3f e2e %.4g graph=%s err=%s launches=%s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["config"].get("cuda_graph"),str(d["config"].get("cuda_graph_error"))[:100],d["gpu_launches"]))
PY
