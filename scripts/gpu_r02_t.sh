#!/bin/bash
# round 2, GPU call T: geometry one batch ahead (second stream inside the step graph) -- parity test, then A/B of the
# bench with / without it and of the SM reserve for the sampling kernels
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests/test_train_step_gpu.py -m gpu -x -q -s > gpurun_out/t_train.log 2>&1; echo "train tests rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/t_train.log)"; grep -E "^FAILED|^ERROR|^losses" gpurun_out/t_train.log | cut -c1-400
run() { tag=$1; shift; flags=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra $flags > gpurun_out/t_bench_$tag.json 2> gpurun_out/t_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0)) $(tail -c 300 gpurun_out/t_bench_$tag.err | tr '\n' ' ')"; }
run plain "--no-ahead" A=1
run ahead "" A=1
run r8 "" PN2_AHEAD_RESERVE=8
run r24 "" PN2_AHEAD_RESERVE=24
run r32 "" PN2_AHEAD_RESERVE=32
run r0 "" PN2_AHEAD_RESERVE=0
run step16 "" PN2_AHEAD_SCOPE=step
run plainb "--no-ahead" A=1
run aheadb "" A=1
python - <<'PY'
import json
for tag in ("plain", "ahead", "r8", "r24", "r32", "r0", "step16", "plainb", "aheadb"):
    try:
        d = json.loads(open("gpurun_out/t_bench_%s.json" % tag).read().strip().splitlines()[-1])
        print("%-7s %.3f ms/step value %.4g e2e %.4g graph %s ahead %s loss %.4f launches %s" % (
            tag, d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["cuda_graph"], bool(d["config"].get("geometry_ahead")),
            d["e2e"]["last_loss"], d["gpu_launches"]))
    except Exception as e:
        print(tag, "parse error", e)
PY
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/t_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/t_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/t_suite.log | cut -c1-300
