#!/bin/bash
# round 2, GPU call L: wgrad with two transform groups; smoke(); per-module parity test; suite; bench; trace; op sweep
mkdir -p gpurun_out
T0=$SECONDS
L=$PWD/open3d-pointnet2-semantic3d_b200/lib
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/l_smoke.log | cut -c1-200)"
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x > gpurun_out/l_gemm.log 2>&1; echo "gemm tests rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/l_gemm.log)"
timeout 200 python -m pytest tests/test_layers_gpu.py -q -s -k identical_inputs 2>&1 | grep -E "per-module|passed|failed|Error" | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/l_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/l_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/l_suite.log | cut -c1-300
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/l_bench_$tag.json 2> gpurun_out/l_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0))"; }
run new PN2_X=1
run new2 PN2_X=1
python - <<'PY'
import json
for tag in ("new", "new2"):
    try:
        d = json.loads(open("gpurun_out/l_bench_%s.json" % tag).read().strip().splitlines()[-1])
        pe = d["roofline"]["per_entry_point"]
        print("%-6s %.3f ms/step e2e %.4g frac %.3f fused %.4f | fwd %.3f dgrad %.3f wgrad %.3f" % (tag, d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fused_chain_model"]["frac"], pe["pn2_linear_fwd"]["ms_per_step"], pe["pn2_linear_dgrad"]["ms_per_step"], pe["pn2_linear_wgrad"]["ms_per_step"]))
    except Exception as e:
        print(tag, "parse error", e)
try:
    d = json.loads(open("gpurun_out/l_bench_new.json").read().strip().splitlines()[-1])
    for r in [r for r in d["linear_calls"] if r["call"] == "wgrad"][:12]:
        print("  %-6s M=%-7d K=%-4d N=%-4d x%.0f  %7.1f us  %6.0f GB/s" % (r["call"], r["M"], r["K"], r["N"], r["calls_per_step"], r["us"], r["GBps"]))
except Exception as e:
    print("table error", e)
PY
PN2_LIB=$L/libpn2_b200_trace.so timeout 120 python scripts/debug_tc_trace.py > gpurun_out/l_trace.log 2>&1; echo "trace rc=$? t=$((SECONDS-T0))"; grep -A6 "wgrad M" gpurun_out/l_trace.log | grep -E "^==|mma issue|mma wait full|mma total" | cut -c1-100
timeout 200 python profiles/op_sweep.py --budget 120 --out gpurun_out/op_sweep_r02.json > gpurun_out/l_sweep.log 2>&1; echo "sweep rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/l_sweep.log | cut -c1-200
