#!/bin/bash
# round 2, GPU call N: BN finalize fused into the GEMM's last CTA; FPS thread-shape A/B (512x16 default, 1024x8, 256x32)
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x > gpurun_out/n_gemm.log 2>&1; echo "gemm tests rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/n_gemm.log)"; grep -E "^FAILED|Error" gpurun_out/n_gemm.log | head -5 | cut -c1-300
STRESS_ITERS=6 timeout 180 python scripts/stress_tc.py > gpurun_out/n_stress.log 2>&1; echo "stress rc=$? $(tail -1 gpurun_out/n_stress.log)"
timeout 100 python scripts/ab_fps.py gpurun_out/fps_512.npy; PN2_FPS_T=1024 timeout 100 python scripts/ab_fps.py gpurun_out/fps_1024.npy; PN2_FPS_T=256 timeout 100 python scripts/ab_fps.py gpurun_out/fps_256.npy
python -c "import numpy as np; a=np.load('gpurun_out/fps_512.npy'); print('identical:', bool((a==np.load('gpurun_out/fps_1024.npy')).all()), bool((a==np.load('gpurun_out/fps_256.npy')).all()))"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/n_suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/n_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/n_suite.log | cut -c1-300
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/n_bench_$tag.json 2> gpurun_out/n_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0))"; }
run a PN2_X=1
run b PN2_X=1
run fps1024 PN2_FPS_T=1024
python - <<'PY'
import json
for tag in ("a", "b", "fps1024"):
    try:
        d = json.loads(open("gpurun_out/n_bench_%s.json" % tag).read().strip().splitlines()[-1])
        pe = d["roofline"]["per_entry_point"]; bd = d["breakdown_ms_per_step"]
        print("%-8s %.3f ms/step e2e %.4g frac %.3f | fwd %.3f dgrad %.3f wgrad %.3f fps %.3f finalize %s launches/step %d" % (tag, d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], pe["pn2_linear_fwd"]["ms_per_step"], pe["pn2_linear_dgrad"]["ms_per_step"], pe["pn2_linear_wgrad"]["ms_per_step"], bd["pn2_fps"]["ms_per_step"], bd.get("pn2_bn_train_finalize"), d["gpu_launches"] // d["steps"]))
    except Exception as e:
        print(tag, "parse error", e)
PY
