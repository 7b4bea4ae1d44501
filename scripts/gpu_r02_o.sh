#!/bin/bash
# round 2, GPU call O: packed fp32x2 FPS, FP4 dgrad column range, wgrad feature-tail A/B
mkdir -p gpurun_out
T0=$SECONDS
timeout 100 python scripts/ab_fps.py gpurun_out/fps_x2.npy
python -c "import numpy as np; print('identical to the scalar kernel of the previous call:', bool((np.load('gpurun_out/fps_x2.npy')==np.load('gpurun_out/fps_512.npy')).all()))"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/o_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/o_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/o_suite.log | cut -c1-300
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/o_bench_$tag.json 2> gpurun_out/o_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0))"; }
run a PN2_X=1
run tail PN2_WGRAD_TAIL=1
run b PN2_X=1
python - <<'PY'
import json
for tag in ("a", "tail", "b"):
    try:
        d = json.loads(open("gpurun_out/o_bench_%s.json" % tag).read().strip().splitlines()[-1])
        pe = d["roofline"]["per_entry_point"]; bd = d["breakdown_ms_per_step"]
        print("%-6s %.3f ms/step e2e %.4g frac %.3f | fwd %.3f dgrad %.3f wgrad %.3f fps %.3f" % (tag, d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], pe["pn2_linear_fwd"]["ms_per_step"], pe["pn2_linear_dgrad"]["ms_per_step"], pe["pn2_linear_wgrad"]["ms_per_step"], bd["pn2_fps"]["ms_per_step"]))
    except Exception as e:
        print(tag, "parse error", e)
for tag in ("a", "tail"):
    try:
        d = json.loads(open("gpurun_out/o_bench_%s.json" % tag).read().strip().splitlines()[-1])
        for r in [r for r in d["linear_calls"] if r["K"] in (131, 259) or (r["call"] == "dgrad" and r["K"] == 128 and r["M"] == 131072)][:8]:
            print("  %-5s %-6s M=%-7d K=%-4d N=%-4d x%.0f  %7.1f us" % (tag, r["call"], r["M"], r["K"], r["N"], r["calls_per_step"], r["us"]))
    except Exception as e:
        print("table error", e)
PY
