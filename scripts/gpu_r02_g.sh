#!/bin/bash
# round 2, GPU call G: same-box A/B of the tc_gemm changes; fused-layer tests; compute-sanitizer memcheck
mkdir -p gpurun_out
T0=$SECONDS
L=$PWD/open3d-pointnet2-semantic3d_b200/lib
timeout 300 python -m pytest tests/test_fused_layers_gpu.py -q > gpurun_out/g_fused.log 2>&1; echo "fused tests rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/g_fused.log)"; grep -E "^FAILED|^ERROR|Error" gpurun_out/g_fused.log | head -10 | cut -c1-300
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/g_bench_$tag.json 2> gpurun_out/g_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0))"; }
run a_xf4      PN2_LIB=$L/libpn2_b200_xf4.so PN2_TC_STACK=0 PN2_TC_EPI_ALT=0
run b_xf8      PN2_TC_STACK=0 PN2_TC_EPI_ALT=0
run c_xf8_alt  PN2_TC_STACK=0
run d_all      PN2_X=1
run a2_xf4     PN2_LIB=$L/libpn2_b200_xf4.so PN2_TC_STACK=0 PN2_TC_EPI_ALT=0
run d2_all     PN2_X=1
python - <<'PY'
import json
for tag in ("a_xf4", "b_xf8", "c_xf8_alt", "d_all", "a2_xf4", "d2_all"):
    try:
        d = json.loads(open("gpurun_out/g_bench_%s.json" % tag).read().strip().splitlines()[-1])
        pe = d["roofline"]["per_entry_point"]
        print("%-10s %.3f ms/step e2e %.4g frac %.3f | fwd %.3f dgrad %.3f wgrad %.3f" % (tag, d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], pe["pn2_linear_fwd"]["ms_per_step"], pe["pn2_linear_dgrad"]["ms_per_step"], pe["pn2_linear_wgrad"]["ms_per_step"]))
    except Exception as e:
        print(tag, "parse error", e)
try:
    d = json.loads(open("gpurun_out/g_bench_d_all.json").read().strip().splitlines()[-1])
    for r in d["linear_calls"][:45]:
        print("  %-6s M=%-7d K=%-4d N=%-4d x%.0f  %7.1f us  %6.0f GB/s" % (r["call"], r["M"], r["K"], r["N"], r["calls_per_step"], r["us"], r["GBps"]))
except Exception as e:
    print("table error", e)
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_ops.py > gpurun_out/g_memcheck.log 2>&1; echo "memcheck rc=$? t=$((SECONDS-T0))"; tail -6 gpurun_out/g_memcheck.log | cut -c1-200
