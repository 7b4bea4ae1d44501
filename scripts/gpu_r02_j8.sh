#!/bin/bash
# round 2, GPU call K (8 GPUs): the scaling line the driver measures at round end
mkdir -p gpurun_out
T0=$SECONDS
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/k_bench$n.json 2> gpurun_out/k_bench$n.err; echo "bench$n rc=$? t=$((SECONDS-T0))"
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/k_bench1.json 2> gpurun_out/k_bench1.err; echo "bench1 rc=$? t=$((SECONDS-T0))"
python - <<'PY'
import json
base = None
for tag in ("k_bench1", "k_bench4", "k_bench8"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % tag).read().strip().splitlines()[-1])
        if base is None: base = d["value"]
        print("%-9s n=%d %.3f ms/step value %.4g eff %.3f e2e %.4g graph %s collective %s" % (tag, d["n_gpus"], d["ms_per_step"], d["value"], d["value"] / (base * d["n_gpus"]), d["e2e"]["value"], d["config"]["cuda_graph"], d.get("collective")))
    except Exception as e:
        print(tag, "parse error", e)
PY
