#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_graph2.py > gpurun_out/debug_graph2.log 2>&1; tail -30 gpurun_out/debug_graph2.log | cut -c1-200
timeout 600 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; cut -c1-110 gpurun_out/bench_gemm.log | head -9
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; tail -3 gpurun_out/bench_graph.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_graph.json"))
print("value %.4g ms/step %.3f e2e %.4g graph=%s err=%s launches=%s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["config"].get("cuda_graph"),str(d["config"].get("cuda_graph_error"))[:200],d["gpu_launches"]))
for k,v in list(d["breakdown_ms_per_step"].items())[:10]: print("   %-32s %.3f ms (%d calls)"%(k,v["ms_per_step"],v["calls_per_step"]))
PY
