#!/bin/bash
# round 2, GPU call R: programmatic dependent launch on every kernel (A/B PN2_PDL), vectorised group_concat / copy_cols /
# dropout, vector reductions + fewer blocks in the skinny wgrad (A/B PN2_SKINNY_BPS)
mkdir -p gpurun_out
T0=$SECONDS
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/r_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/r_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/r_suite.log | cut -c1-300
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r_bench_$tag.json 2> gpurun_out/r_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0))"; }
run pdl1 PN2_PDL=1
run pdl0 PN2_PDL=0
run pdl1b PN2_PDL=1
run pdl0b PN2_PDL=0
run bps6 PN2_PDL=1 PN2_SKINNY_BPS=6
run bps2 PN2_PDL=1 PN2_SKINNY_BPS=2
python - <<'PY'
import json
for tag in ("pdl1", "pdl0", "pdl1b", "pdl0b", "bps6", "bps2"):
    try:
        d = json.loads(open("gpurun_out/r_bench_%s.json" % tag).read().strip().splitlines()[-1])
        bd = d["breakdown_ms_per_step"]
        sk = [r for r in d["linear_calls"] if r["call"] == "wgrad" and r["N"] == 9]
        print("%-6s %.3f ms/step e2e %.4g | concat %.3f copy_cols %.3f dropout %.3f wgrad %.3f skinny %.1f us | eager sum %.3f" % (
            tag, d["ms_per_step"], d["e2e"]["value"], bd["pn2_group_concat_ld"]["ms_per_step"], bd["pn2_copy_cols"]["ms_per_step"],
            bd["pn2_dropout"]["ms_per_step"], bd["pn2_linear_wgrad"]["ms_per_step"], sk[0]["us"] if sk else -1,
            sum(v["ms_per_step"] for v in bd.values())))
    except Exception as e:
        print(tag, "parse error", e)
PY
