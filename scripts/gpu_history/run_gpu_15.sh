#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_gemm_gpu.py -q -m gpu > gpurun_out/layers_tc.log 2>&1
echo "layers+gemm rc=$?"; grep -E "passed|failed|^FAILED|AssertionError" gpurun_out/layers_tc.log | cut -c1-300 | head -20
for mode in graph nograph; do
  if [ $mode = nograph ]; then EXTRA="--no-graph --no-cpu-baseline"; else EXTRA=""; fi
  timeout 600 python bench.py --steps 10 --warmup 3 $EXTRA > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; tail -3 gpurun_out/bench_$mode.err
done
python - <<'PY'
import json
for f in ("bench_graph","bench_nograph"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "value %.4g ms/step %.3f e2e %.4g graph=%s launches=%s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["config"].get("cuda_graph"),d["gpu_launches"]), d.get("clocks"))
        for k,v in list(d["breakdown_ms_per_step"].items())[:10]: print("   %-32s %.3f ms (%d calls)"%(k,v["ms_per_step"],v["calls_per_step"]))
    except Exception as e: print(f, "ERR", e)
PY
