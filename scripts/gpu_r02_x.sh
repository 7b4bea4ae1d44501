#!/bin/bash
# round 2, GPU call X: state after geometry-ahead + side-stream weight gradients -- suite, smoke, full bench line, reference arm, ncu launch list + captures
mkdir -p gpurun_out
T0=$SECONDS
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/x_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/x_smoke.log | cut -c1-200)"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/x_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/x_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/x_suite.log | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; echo "bench rc=$? t=$((SECONDS-T0))"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/x_bench.json").read().strip().splitlines()[-1])
    pe = d["roofline"]["per_entry_point"]
    print("%.3f ms/step value %.4g e2e %.4g frac %.3f fused %.4f | fwd %.3f dgrad %.3f wgrad %.3f launches/step %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fused_chain_model"]["frac"], pe["pn2_linear_fwd"]["ms_per_step"], pe["pn2_linear_dgrad"]["ms_per_step"], pe["pn2_linear_wgrad"]["ms_per_step"], d["gpu_launches"] // d["steps"]))
    print("cpu", d["cpu_baseline"]["value"], "config1", d["config1"].get("gpu_graph_ms"), d["config1"].get("cpu_ms"), "cfeat6", d["cfeat6"].get("ms_per_step"))
    for k, v in list(d["breakdown_ms_per_step"].items())[:26]: print("  %-30s %.3f ms x%d" % (k, v["ms_per_step"], v["calls_per_step"]))
except Exception as e:
    print("parse error", e)
PY
P="python scripts/profile_step.py"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv $P > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 19 -c 2 -o gpurun_out/tcgemm_r02 -f $P > gpurun_out/ncu_tc.log 2>&1; tail -1 gpurun_out/ncu_tc.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -c 1 -o gpurun_out/tcwgrad_r02 -f $P > gpurun_out/ncu_wg.log 2>&1; tail -1 gpurun_out/ncu_wg.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:fps_pruned_kernel -c 1 -o gpurun_out/fps_r02 -f $P > gpurun_out/ncu_fps.log 2>&1; tail -1 gpurun_out/ncu_fps.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:bn_bwd_reduce_v4 -c 1 -o gpurun_out/bnreduce_r02 -f $P > gpurun_out/ncu_bn.log 2>&1; tail -1 gpurun_out/ncu_bn.log
echo "done t=$((SECONDS-T0))"
