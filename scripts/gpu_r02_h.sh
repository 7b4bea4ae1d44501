#!/bin/bash
# round 2, GPU call H: warp-uniform MMA issue loop (no R2UR waterfall): stress, GEMM tests, suite, bench x2, trace
mkdir -p gpurun_out
T0=$SECONDS
L=$PWD/open3d-pointnet2-semantic3d_b200/lib
STRESS_ITERS=8 timeout 180 python scripts/stress_tc.py > gpurun_out/h_stress.log 2>&1; echo "stress rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/h_stress.log)"; grep -v " 0 / " gpurun_out/h_stress.log | head -10 | cut -c1-200
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x > gpurun_out/h_gemm.log 2>&1; echo "gemm tests rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/h_gemm.log)"
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/h_bench_$tag.json 2> gpurun_out/h_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0))"; }
run d_all      PN2_X=1
run d2_all     PN2_X=1
run nostack    PN2_TC_STACK=0
python - <<'PY'
import json
for tag in ("d_all", "d2_all", "nostack"):
    try:
        d = json.loads(open("gpurun_out/h_bench_%s.json" % tag).read().strip().splitlines()[-1])
        pe = d["roofline"]["per_entry_point"]
        print("%-10s %.3f ms/step e2e %.4g frac %.3f | fwd %.3f dgrad %.3f wgrad %.3f" % (tag, d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], pe["pn2_linear_fwd"]["ms_per_step"], pe["pn2_linear_dgrad"]["ms_per_step"], pe["pn2_linear_wgrad"]["ms_per_step"]))
    except Exception as e:
        print(tag, "parse error", e)
try:
    d = json.loads(open("gpurun_out/h_bench_d_all.json").read().strip().splitlines()[-1])
    for r in d["linear_calls"][:30]:
        print("  %-6s M=%-7d K=%-4d N=%-4d x%.0f  %7.1f us  %6.0f GB/s" % (r["call"], r["M"], r["K"], r["N"], r["calls_per_step"], r["us"], r["GBps"]))
    for k, v in list(d["breakdown_ms_per_step"].items())[:24]: print("  %-28s %.3f ms x%d" % (k, v["ms_per_step"], v["calls_per_step"]))
except Exception as e:
    print("table error", e)
PY
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/h_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/h_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/h_suite.log | cut -c1-300
PN2_LIB=$L/libpn2_b200_trace.so timeout 120 python scripts/debug_tc_trace.py > gpurun_out/h_trace.log 2>&1; echo "trace rc=$? t=$((SECONDS-T0))"; grep -E "^==|mma issue|mma wait full|mma total|epi process|prod store|prod load" gpurun_out/h_trace.log | cut -c1-100
