#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -k "bn_backward" > gpurun_out/bn_test.log 2>&1
echo "bn test rc=$?"; grep -E "passed|failed|^FAILED|Max abs|Mismatch" gpurun_out/bn_test.log | head
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; tail -3 gpurun_out/bench_graph.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_graph.json"))
print("value %.4g ms/step %.3f e2e %.4g graph=%s err=%s launches=%s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["config"].get("cuda_graph"),d["config"].get("cuda_graph_error"),d["gpu_launches"]))
for k,v in list(d["breakdown_ms_per_step"].items())[:8]: print("   %-32s %.3f ms (%d calls)"%(k,v["ms_per_step"],v["calls_per_step"]))
PY
