#!/bin/bash
# Last GPU call of round 1 (about 7 GPU-minutes were left): new-op parity tests, ncu captures of the
# index/gather ops at config-2 sizes, the config-5 sweep and the config-3 MSG layer.
mkdir -p gpurun_out
T0=$SECONDS
timeout 170 python -m pytest tests/test_ops_gpu.py -q -x -k "prob_sample or interpolate_label" > gpurun_out/newops_pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))"; tail -4 gpurun_out/newops_pytest.log
timeout 100 ncu --set full --clock-control none --import-source on \
  -k 'regex:ball_query|three_nn_kernel|group_point_kernel|three_interpolate_kernel' -c 4 \
  -o gpurun_out/ops_r01b -f python profiles/op_sweep.py --only cfg2 > gpurun_out/ncu_ops.log 2>&1; echo "ncu rc=$? t=$((SECONDS-T0))"; tail -2 gpurun_out/ncu_ops.log
timeout 60 python profiles/op_sweep.py --only msg --out gpurun_out/op_msg.json > gpurun_out/op_msg.log 2>&1; echo "msg rc=$? t=$((SECONDS-T0))"; tail -2 gpurun_out/op_msg.log | cut -c1-400
timeout 150 python profiles/op_sweep.py --budget 45 > gpurun_out/op_sweep.log 2>&1; echo "sweep rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/op_sweep.log | cut -c1-300
ls -la gpurun_out | head -20
