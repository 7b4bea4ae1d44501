#!/bin/bash
# round 2, GPU call C: whole suite (strict gradient checks, new trainer tests), bench, role-cycle trace of the tcgen05 kernels
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/c_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/c_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/c_suite.log | cut -c1-300
timeout 120 python -m pytest tests/test_train_step_gpu.py -q -s -k graph_replay 2>&1 | grep -i "noise" | cut -c1-600
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/c_bench.err | cut -c1-1000
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c_bench.json").read().strip().splitlines()[-1])
    print("ms/step %.3f value %.4g e2e %.4g ratio %.3f graph %s launches %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["config"]["cuda_graph"], d["gpu_launches"]))
    print("err", d["config"].get("cuda_graph_error"))
    print("roofline", d["roofline"]["frac"], d["roofline"].get("per_entry_point"))
    for k, v in list(d["breakdown_ms_per_step"].items())[:12]: print("  %-28s %.3f ms x%d" % (k, v["ms_per_step"], v["calls_per_step"]))
except Exception as e:
    print("parse error", e)
PY
PN2_LIB=$PWD/open3d-pointnet2-semantic3d_b200/lib/libpn2_b200_trace.so timeout 120 python scripts/debug_tc_trace.py > gpurun_out/c_trace.log 2>&1; echo "trace rc=$? t=$((SECONDS-T0))"; cat gpurun_out/c_trace.log | cut -c1-200
