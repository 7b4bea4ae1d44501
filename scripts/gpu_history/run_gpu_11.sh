#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gemm_gpu.py -q -m gpu > gpurun_out/gemm_test.log 2>&1
echo "gemm test rc=$?"; grep -E "passed|failed|^FAILED|Max abs|Mismatch" gpurun_out/gemm_test.log | head -30
timeout 600 python -m pytest tests/test_layers_gpu.py -q -m gpu > gpurun_out/layers_tc.log 2>&1
echo "layers tc rc=$?"; grep -E "passed|failed|^FAILED|AssertionError" gpurun_out/layers_tc.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; tail -2 gpurun_out/bench_tc.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_tc.json"))
print("value %.4g ms/step %.3f e2e %.4g"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d.get("clocks"))
for k,v in list(d["breakdown_ms_per_step"].items())[:16]: print("   %-32s %.3f ms (%d calls)"%(k,v["ms_per_step"],v["calls_per_step"]))
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python scripts/profile_step.py > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
