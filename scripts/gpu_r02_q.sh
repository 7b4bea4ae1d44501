#!/bin/bash
# round 2, GPU call Q: vector reductions in the scatter-add gradient kernels; BN-backward reduce blocks-per-SM A/B
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/q_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/q_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/q_suite.log | cut -c1-300
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/q_bench_$tag.json 2> gpurun_out/q_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0))"; }
run bps2 PN2_BNRED_BPS=2
run bps3 PN2_BNRED_BPS=3
run bps4 PN2_BNRED_BPS=4
run bps2b PN2_BNRED_BPS=2
python - <<'PY'
import json
for tag in ("bps2", "bps3", "bps4", "bps2b"):
    try:
        d = json.loads(open("gpurun_out/q_bench_%s.json" % tag).read().strip().splitlines()[-1])
        bd = d["breakdown_ms_per_step"]
        print("%-6s %.3f ms/step e2e %.4g | bn_reduce %.3f bn_apply %.3f interp_grad %.3f concat_grad %.3f" % (tag, d["ms_per_step"], d["e2e"]["value"], bd["pn2_bn_bwd_reduce"]["ms_per_step"], bd["pn2_bn_bwd_apply"]["ms_per_step"], bd["pn2_three_interpolate_grad_ld"]["ms_per_step"], bd["pn2_group_concat_grad"]["ms_per_step"]))
    except Exception as e:
        print(tag, "parse error", e)
PY
