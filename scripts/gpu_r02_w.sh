#!/bin/bash
# round 2, GPU call W: side-stream weight gradients -- stream priority, full device for the last wgrad, SM split sweep
mkdir -p gpurun_out
T0=$SECONDS
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/w_bench_$tag.json 2> gpurun_out/w_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0)) $(tail -c 300 gpurun_out/w_bench_$tag.err | tr '\n' ' ')"; }
run p0_w0 PN2_MAIN_PRIO=0 PN2_WGRAD_SMS=0
run p1_w0 PN2_MAIN_PRIO=1 PN2_WGRAD_SMS=0
run p0_w64t PN2_MAIN_PRIO=0 PN2_WGRAD_SMS=64
run p1_w64t PN2_MAIN_PRIO=1 PN2_WGRAD_SMS=64
run p1_w64nt PN2_MAIN_PRIO=1 PN2_WGRAD_SMS=64 PN2_WGRAD_TAIL_FULL=0
run p1_w48t PN2_MAIN_PRIO=1 PN2_WGRAD_SMS=48
run p1_w56t PN2_MAIN_PRIO=1 PN2_WGRAD_SMS=56
run p1_w72t PN2_MAIN_PRIO=1 PN2_WGRAD_SMS=72
run p1_w80t PN2_MAIN_PRIO=1 PN2_WGRAD_SMS=80
run p1_w148 PN2_MAIN_PRIO=1 PN2_WGRAD_SMS=148
run p1_w64t_r24 PN2_MAIN_PRIO=1 PN2_WGRAD_SMS=64 PN2_AHEAD_RESERVE=24
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/w_bench_*.json")):
    tag = f.split("w_bench_")[1][:-5]
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-12s %.3f ms/step value %.4g e2e %.4g graph %s loss %.4f" % (
            tag, d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["cuda_graph"], d["e2e"]["last_loss"]))
    except Exception as e:
        print(tag, "parse error", e)
PY
timeout 600 python -m pytest tests/test_train_step_gpu.py -m gpu -x -q > gpurun_out/w_train.log 2>&1; echo "train tests rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/w_train.log)"
