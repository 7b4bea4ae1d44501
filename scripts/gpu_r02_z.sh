#!/bin/bash
# round 2, GPU call Z: GEMM prologue trims (bias load behind an epilogue-only barrier, tensormap prefetch), weight images
# prepared on the side stream, rotating batches in the bench
mkdir -p gpurun_out
T0=$SECONDS
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/z_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/z_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/z_suite.log | cut -c1-300
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/z_bench_$tag.json 2> gpurun_out/z_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0)) $(tail -c 300 gpurun_out/z_bench_$tag.err | tr '\n' ' ')"; }
run side1 PN2_PREP_SIDE=1
run side0 PN2_PREP_SIDE=0
run side1b PN2_PREP_SIDE=1
run side0b PN2_PREP_SIDE=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/z_bench_*.json")):
    tag = f.split("z_bench_")[1][:-5]
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-8s %.3f ms/step value %.4g e2e %.4g graph %s loss %.4f" % (
            tag, d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["cuda_graph"], d["e2e"]["last_loss"]))
    except Exception as e:
        print(tag, "parse error", e)
PY
