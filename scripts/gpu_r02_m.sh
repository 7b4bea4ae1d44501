#!/bin/bash
# round 2, GPU call M: after reverting the wgrad two-group change: tests, FPS 512x16 A/B, MSG layer timing, full bench
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x > gpurun_out/m_gemm.log 2>&1; echo "gemm tests rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/m_gemm.log)"
timeout 300 python -m pytest tests/test_layers_gpu.py tests/test_train_step_gpu.py -q -s -k "identical_inputs or graph_replay" 2>&1 | grep -E "per-module|noise|passed|failed|Error" | cut -c1-500
timeout 100 python scripts/ab_fps.py gpurun_out/fps_a.npy; PN2_FPS_T512=1 timeout 100 python scripts/ab_fps.py gpurun_out/fps_b.npy
python -c "import numpy as np; a=np.load('gpurun_out/fps_a.npy'); b=np.load('gpurun_out/fps_b.npy'); print('fps variants identical:', bool((a==b).all()))"
timeout 200 python profiles/op_sweep.py --only msg --out gpurun_out/op_msg_r02.json > gpurun_out/m_msg.log 2>&1; echo "msg rc=$? t=$((SECONDS-T0))"; tail -2 gpurun_out/m_msg.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/m_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/m_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/m_suite.log | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$? t=$((SECONDS-T0))"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/m_bench_ref.json 2> gpurun_out/m_bench_ref.err; echo "bench ref rc=$? t=$((SECONDS-T0))"; cut -c1-300 gpurun_out/m_bench_ref.json
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/m_bench.json").read().strip().splitlines()[-1])
    pe = d["roofline"]["per_entry_point"]
    print("%.3f ms/step value %.4g e2e %.4g frac %.3f fused %.4f | fwd %.3f dgrad %.3f wgrad %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fused_chain_model"]["frac"], pe["pn2_linear_fwd"]["ms_per_step"], pe["pn2_linear_dgrad"]["ms_per_step"], pe["pn2_linear_wgrad"]["ms_per_step"]))
    print("cpu", d["cpu_baseline"]["value"], "config1", d["config1"].get("gpu_graph_ms"), d["config1"].get("cpu_ms"), "cfeat6", d["cfeat6"].get("ms_per_step"))
except Exception as e:
    print("parse error", e)
PY
