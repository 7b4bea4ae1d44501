#!/bin/bash
# round 2, GPU call E: tc_gemm with two transform groups + alternate-tile epilogue: stress, GEMM tests, suite, bench, trace
mkdir -p gpurun_out
T0=$SECONDS
STRESS_ITERS=12 timeout 180 python scripts/stress_tc.py > gpurun_out/e_stress.log 2>&1; echo "stress rc=$? t=$((SECONDS-T0))"; tail -14 gpurun_out/e_stress.log | cut -c1-200
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x > gpurun_out/e_gemm.log 2>&1; echo "gemm tests rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/e_gemm.log)"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/e_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/e_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/e_suite.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/e_bench.err | cut -c1-600
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/e_bench.json").read().strip().splitlines()[-1])
    print("ms/step %.3f value %.4g e2e %.4g ratio %.3f graph %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["config"]["cuda_graph"]))
    print("roofline", d["roofline"]["frac"], d["roofline"].get("per_entry_point"))
    for k, v in list(d["breakdown_ms_per_step"].items())[:10]: print("  %-28s %.3f ms x%d" % (k, v["ms_per_step"], v["calls_per_step"]))
except Exception as e:
    print("parse error", e)
PY
PN2_LIB=$PWD/open3d-pointnet2-semantic3d_b200/lib/libpn2_b200_trace.so timeout 120 python scripts/debug_tc_trace.py > gpurun_out/e_trace.log 2>&1; echo "trace rc=$? t=$((SECONDS-T0))"; head -60 gpurun_out/e_trace.log | cut -c1-120
