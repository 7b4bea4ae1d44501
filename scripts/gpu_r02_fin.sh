#!/bin/bash
# round 2, final GPU call: smoke, suite, full bench line, reference arm, ncu launch list + GEMM captures of HEAD
mkdir -p gpurun_out
T0=$SECONDS
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/fin_smoke.log | cut -c1-200)"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/fin_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/fin_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/fin_suite.log | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; echo "bench rc=$? t=$((SECONDS-T0))"
timeout 400 python bench.py > gpurun_out/fin_bench_default.json 2> gpurun_out/fin_bench_default.err; echo "bench(default args) rc=$? t=$((SECONDS-T0))"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fin_ref.json 2> gpurun_out/fin_ref.err; echo "reference arm rc=$? t=$((SECONDS-T0)) $(cut -c1-200 gpurun_out/fin_ref.json)"
python - <<'PY'
import json
for f in ("fin_bench", "fin_bench_default"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        pe = d["roofline"]["per_entry_point"]
        print("%s: %.3f ms/step value %.4g e2e %.4g frac %.3f | fwd %.3f dgrad %.3f wgrad %.3f launches/step %d cpu %.0f cfeat6 %s" % (f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], pe["pn2_linear_fwd"]["ms_per_step"], pe["pn2_linear_dgrad"]["ms_per_step"], pe["pn2_linear_wgrad"]["ms_per_step"], d["gpu_launches"] // d["steps"], d["cpu_baseline"]["value"], d["cfeat6"].get("ms_per_step")))
    except Exception as e:
        print(f, "parse error", e)
PY
P="python scripts/profile_step.py"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv $P > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 19 -c 2 -o gpurun_out/tcgemm_r02 -f $P > gpurun_out/ncu_tc.log 2>&1; tail -1 gpurun_out/ncu_tc.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -c 1 -o gpurun_out/tcwgrad_r02 -f $P > gpurun_out/ncu_wg.log 2>&1; tail -1 gpurun_out/ncu_wg.log
echo "done t=$((SECONDS-T0))"
