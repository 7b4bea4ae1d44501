#!/bin/bash
# round 2, GPU call V: weight gradients on a second stream during the backward pass -- parity tests, then a sweep of the
# SM split in the bench (geometry-ahead on)
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests/test_train_step_gpu.py -m gpu -x -q -s > gpurun_out/v_train.log 2>&1; echo "train tests rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/v_train.log)"; grep -E "^FAILED|^ERROR|^losses|^E  " gpurun_out/v_train.log | cut -c1-500
run() { tag=$1; shift; flags=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra $flags > gpurun_out/v_bench_$tag.json 2> gpurun_out/v_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0)) $(tail -c 300 gpurun_out/v_bench_$tag.err | tr '\n' ' ')"; }
run w0 "" PN2_WGRAD_SMS=0
run w32 "" PN2_WGRAD_SMS=32
run w48 "" PN2_WGRAD_SMS=48
run w64 "" PN2_WGRAD_SMS=64
run w80 "" PN2_WGRAD_SMS=80
run w100 "" PN2_WGRAD_SMS=100
run w148 "" PN2_WGRAD_SMS=148
run w0b "" PN2_WGRAD_SMS=0
python - <<'PY'
import json
for tag in ("w0", "w32", "w48", "w64", "w80", "w100", "w148", "w0b"):
    try:
        d = json.loads(open("gpurun_out/v_bench_%s.json" % tag).read().strip().splitlines()[-1])
        print("%-7s %.3f ms/step value %.4g e2e %.4g graph %s ahead %s loss %.4f" % (
            tag, d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["cuda_graph"], bool(d["config"].get("geometry_ahead")),
            d["e2e"]["last_loss"]))
    except Exception as e:
        print(tag, "parse error", e)
PY
