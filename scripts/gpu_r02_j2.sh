#!/bin/bash
# round 2, GPU call J (2 GPUs): NCCL data-parallel correctness test + 2-GPU bench line
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_train_step_gpu.py -q -k data_parallel > gpurun_out/j_dp.log 2>&1; echo "dp test rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/j_dp.log)"; grep -E "DP_OK|Error|FAILED" gpurun_out/j_dp.log | head -6 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/j_bench2.json 2> gpurun_out/j_bench2.err; echo "bench2 rc=$? t=$((SECONDS-T0))"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/j_bench1.json 2> gpurun_out/j_bench1.err; echo "bench1 rc=$? t=$((SECONDS-T0))"
python - <<'PY'
import json
for tag in ("j_bench1", "j_bench2"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % tag).read().strip().splitlines()[-1])
        print("%-9s n=%d %.3f ms/step value %.4g e2e %.4g graph %s collective %s" % (tag, d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["cuda_graph"], d.get("collective")))
    except Exception as e:
        print(tag, "parse error", e)
PY
