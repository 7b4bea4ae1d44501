#!/bin/bash
# Round-1 last call: the thread-block-cluster FPS kernel -- parity tests, then timing at the config-5 sizes.
mkdir -p gpurun_out
T0=$SECONDS
timeout 90 python -m pytest tests/test_ops_gpu.py -q -x -k "fps_cluster" > gpurun_out/cluster_pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))"; tail -6 gpurun_out/cluster_pytest.log | cut -c1-300
timeout 80 python profiles/op_sweep.py --only fps_cluster --budget 40 --out gpurun_out/op_fps_cluster.json > gpurun_out/op_fps_cluster.log 2>&1; echo "sweep rc=$? t=$((SECONDS-T0))"; tail -7 gpurun_out/op_fps_cluster.log | cut -c1-420
