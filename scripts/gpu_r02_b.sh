#!/bin/bash
# round 2, GPU call B: new single-stream Trainer (graph capture), bench, A/B of the promoted-kernel candidates
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_train_step_gpu.py -x -q -s > gpurun_out/b_train.log 2>&1; echo "train tests rc=$? t=$((SECONDS-T0))"; tail -15 gpurun_out/b_train.log | cut -c1-600
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$? t=$((SECONDS-T0))"; tail -5 gpurun_out/b_bench.err | cut -c1-1500
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/b_bench.json").read().strip().splitlines()[-1])
    print("ms/step %.3f value %.4g e2e %.4g ratio %.3f graph %s launches %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["config"]["cuda_graph"], d["gpu_launches"]))
    print("err", d["config"].get("cuda_graph_error"))
    print("roofline", d["roofline"]["frac"], d["roofline"].get("per_entry_point"))
    print("cpu", d["cpu_baseline"])
    print("config1", d.get("config1")); print("cfeat6", d.get("cfeat6"))
    for k, v in list(d["breakdown_ms_per_step"].items())[:14]: print("  %-28s %.3f ms x%d" % (k, v["ms_per_step"], v["calls_per_step"]))
except Exception as e:
    print("parse error", e)
PY
timeout 200 python scripts/ab_ops.py > gpurun_out/ab_ops.log 2>&1; echo "ab rc=$? t=$((SECONDS-T0))"; cat gpurun_out/ab_ops.log | cut -c1-250
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/b_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/b_suite.log)"
PN2_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_experimental_gpu.py -x -q > gpurun_out/b_exp.log 2>&1; echo "exp rc=$? $(tail -1 gpurun_out/b_exp.log)"
