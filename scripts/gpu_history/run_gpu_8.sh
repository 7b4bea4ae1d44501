#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gemm_gpu.py -q -m gpu > gpurun_out/gemm_test.log 2>&1
echo "gemm test rc=$?"; grep -E "passed|failed|^FAILED|Max abs|Mismatch" gpurun_out/gemm_test.log | head -30
PN2_GEMM_MODE=0 timeout 600 python -m pytest tests/test_layers_gpu.py -q -m gpu > gpurun_out/layers_mode0.log 2>&1
echo "layers mode0 rc=$?"; grep -E "passed|failed|^FAILED|AssertionError" gpurun_out/layers_mode0.log | head -20
PN2_GEMM_MODE=-1 timeout 600 python -m pytest tests/test_layers_gpu.py -q -m gpu > gpurun_out/layers_tc.log 2>&1
echo "layers tc rc=$?"; grep -E "passed|failed|^FAILED|AssertionError" gpurun_out/layers_tc.log | head -20
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
PN2_GEMM_MODE=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mode0.json 2> gpurun_out/bench_mode0.err; tail -2 gpurun_out/bench_mode0.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; tail -2 gpurun_out/bench_tc.err
python - <<'PY'
import json
for f in ("bench_mode0","bench_tc"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "value %.4g ms/step %.3f e2e %.4g"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), d.get("cpu_baseline"))
        for k,v in list(d["breakdown_ms_per_step"].items())[:12]: print("   %-32s %.3f ms (%d calls)"%(k,v["ms_per_step"],v["calls_per_step"]))
    except Exception as e: print(f, "ERR", e)
PY
