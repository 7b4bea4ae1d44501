#!/bin/bash
# Suggested FIRST GPU call of round 2 (see NOTES.md): verify the experimental kernels written blind at the end of
# round 1, re-run the whole suite, and A/B the candidates on the real training step.  ~5 min of run time.
#   gpurun --timeout 600 -- 'bash scripts/gpu_r02_first.sh'
mkdir -p gpurun_out
T0=$SECONDS
# 1. experimental kernels, one pytest process per group so that a hang (wrong handshake) costs only its group
for grp in three_nn_filtered ball_query_grid knn_vote_filtered fps_cluster_mb geometry_prefetch two_training_steps; do
  PN2_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_experimental_gpu.py tests/test_train_step_gpu.py -q -x -k "$grp" \
      > gpurun_out/exp_$grp.log 2>&1
  echo "exp $grp rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/exp_$grp.log | cut -c1-160)"
done
# 2. the whole GPU suite (default paths)
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/suite.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/suite.log)"
# 3. bench A/B on the real step
timeout 150 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err; echo "bench base rc=$? t=$((SECONDS-T0))"
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prefetch > gpurun_out/bench_prefetch.json 2> gpurun_out/bench_prefetch.err; echo "bench prefetch rc=$? t=$((SECONDS-T0))"
PN2_THREE_NN_FILTER=1 PN2_BALL_GRID=1 timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gate_grid.json 2> gpurun_out/bench_gate_grid.err; echo "bench gate+grid rc=$? t=$((SECONDS-T0))"
python - <<'PY'
import json
for tag in ("base", "prefetch", "gate_grid"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % tag).read().strip().splitlines()[-1])
        print("%-10s %.3f ms/step  value %.3g  e2e %.3g  graph %s  err %s" % (
            tag, d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"].get("cuda_graph"),
            d["config"].get("cuda_graph_error")))
    except Exception as e:
        print(tag, "parse error", e)
PY
# 4. op-level timing of the candidates (config-5 sizes)
timeout 120 python profiles/op_sweep.py --only ball_grid --budget 60 --out gpurun_out/op_ball_grid.json > gpurun_out/op_ball_grid.log 2>&1; echo "ball_grid rc=$? t=$((SECONDS-T0))"; tail -4 gpurun_out/op_ball_grid.log | cut -c1-300
timeout 120 python profiles/op_sweep.py --only fps_cluster --fps-entry pn2_fps_cluster_mb --budget 40 --out gpurun_out/op_fps_mb.json > gpurun_out/op_fps_mb.log 2>&1; echo "fps_mb rc=$? t=$((SECONDS-T0))"; tail -4 gpurun_out/op_fps_mb.log | cut -c1-300
