#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/debug_wgrad.py > gpurun_out/debug_wgrad.log 2>&1; head -12 gpurun_out/debug_wgrad.log
timeout 400 python -m pytest tests/test_gemm_gpu.py -q -m gpu > gpurun_out/gemm_test.log 2>&1
echo "gemm test rc=$?"; grep -E "passed|failed|^FAILED|Max abs|Mismatch" gpurun_out/gemm_test.log | head -30
timeout 600 python -m pytest tests/test_layers_gpu.py -q -m gpu > gpurun_out/layers_tc.log 2>&1
echo "layers tc rc=$?"; grep -E "passed|failed|^FAILED|AssertionError" gpurun_out/layers_tc.log | cut -c1-400 | head -20
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log | cut -c1-400
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; tail -2 gpurun_out/bench_tc.err
python - <<'PY'
import json
for f in ("bench_tc",):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "value %.4g ms/step %.3f e2e %.4g"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d.get("cpu_baseline"), d.get("clocks"))
        for k,v in list(d["breakdown_ms_per_step"].items())[:14]: print("   %-32s %.3f ms (%d calls)"%(k,v["ms_per_step"],v["calls_per_step"]))
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python scripts/profile_step.py > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:fps_reg_kernel -c 1 -o gpurun_out/fps_r01 python scripts/profile_step.py > gpurun_out/ncu_fps.log 2>&1; tail -2 gpurun_out/ncu_fps.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 2 -c 2 -o gpurun_out/tcgemm_r01 python scripts/profile_step.py > gpurun_out/ncu_tc.log 2>&1; tail -2 gpurun_out/ncu_tc.log
ls -la gpurun_out | tail -12
