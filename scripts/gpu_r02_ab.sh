#!/bin/bash
# round 2, GPU call AB: late trigger now in the default build; BN-backward reduce blocks per SM under overlap (2 vs 3), 3 runs each
mkdir -p gpurun_out
T0=$SECONDS
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/ab_bench_$tag.json 2> gpurun_out/ab_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0)) $(tail -c 300 gpurun_out/ab_bench_$tag.err | tr '\n' ' ')"; }
for r in a b c; do
run bps2$r PN2_BNRED_BPS=2
run bps3$r PN2_BNRED_BPS=3
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_bench_*.json")):
    tag = f.split("ab_bench_")[1][:-5]
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-8s %.3f ms/step value %.4g e2e %.4g" % (tag, d["ms_per_step"], d["value"], d["e2e"]["value"]))
    except Exception as e:
        print(tag, "parse error", e)
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/ab_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/ab_suite.log | cut -c1-300
