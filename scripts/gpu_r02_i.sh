#!/bin/bash
# round 2, GPU call I: skinny wgrad v4, full bench line, ncu launch list + full captures, racecheck
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x > gpurun_out/i_gemm.log 2>&1; echo "gemm tests rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/i_gemm.log)"
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$? t=$((SECONDS-T0))"; tail -2 gpurun_out/i_bench.err | cut -c1-400
PN2_WGRAD_SKINNY_V1=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/i_bench_skinny1.json 2>/dev/null; echo "bench skinny-v1 rc=$? t=$((SECONDS-T0))"
python - <<'PY'
import json
for tag in ("i_bench", "i_bench_skinny1"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % tag).read().strip().splitlines()[-1])
        pe = d["roofline"]["per_entry_point"]
        print("%-16s %.3f ms/step value %.4g e2e %.4g frac %.3f | fwd %.3f dgrad %.3f wgrad %.3f" % (tag, d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], pe["pn2_linear_fwd"]["ms_per_step"], pe["pn2_linear_dgrad"]["ms_per_step"], pe["pn2_linear_wgrad"]["ms_per_step"]))
        if tag == "i_bench":
            print("cpu", d["cpu_baseline"]); print("config1", d["config1"]); print("cfeat6", d["cfeat6"]); print("clocks", d["clocks"])
    except Exception as e:
        print(tag, "parse error", e)
PY
P="python scripts/profile_step.py"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv $P > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 19 -c 2 -o gpurun_out/tcgemm_r02 -f $P > gpurun_out/ncu_tc.log 2>&1; tail -1 gpurun_out/ncu_tc.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -c 1 -o gpurun_out/tcwgrad_r02 -f $P > gpurun_out/ncu_wg.log 2>&1; tail -1 gpurun_out/ncu_wg.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ball_query_grid_kernel -c 1 -o gpurun_out/ballgrid_r02 -f $P > gpurun_out/ncu_bg.log 2>&1; tail -1 gpurun_out/ncu_bg.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:three_nn_kernel -c 1 -o gpurun_out/threenn_r02 -f $P > gpurun_out/ncu_nn.log 2>&1; tail -1 gpurun_out/ncu_nn.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:fps_pruned_kernel -c 1 -o gpurun_out/fps_r02 -f $P > gpurun_out/ncu_fps.log 2>&1; tail -1 gpurun_out/ncu_fps.log
ls -la gpurun_out/*_r02.ncu-rep
echo "ncu done t=$((SECONDS-T0))"
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_ops.py fps ball knn interp feed > gpurun_out/i_racecheck_index.log 2>&1; echo "racecheck(index ops) rc=$? t=$((SECONDS-T0))"; tail -4 gpurun_out/i_racecheck_index.log | cut -c1-200
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_ops.py gemm > gpurun_out/i_racecheck_gemm.log 2>&1; echo "racecheck(gemm) rc=$? t=$((SECONDS-T0))"; tail -6 gpurun_out/i_racecheck_gemm.log | cut -c1-200
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python scripts/sanitize_ops.py gemm fps ball knn > gpurun_out/i_synccheck.log 2>&1; echo "synccheck rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/i_synccheck.log | cut -c1-200
