#!/bin/bash
mkdir -p gpurun_out
P="python scripts/profile_step.py"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv $P > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 19 -c 2 -o gpurun_out/tcgemm_r01 -f $P > gpurun_out/ncu_tc.log 2>&1; tail -1 gpurun_out/ncu_tc.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -c 1 -o gpurun_out/tcwgrad_r01 -f $P > gpurun_out/ncu_wg.log 2>&1; tail -1 gpurun_out/ncu_wg.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:fps_pruned_kernel -c 1 -o gpurun_out/fps_r01 -f $P > gpurun_out/ncu_fps.log 2>&1; tail -1 gpurun_out/ncu_fps.log
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:bn_bwd_reduce_v4 -c 1 -o gpurun_out/bnreduce_r01 -f $P > gpurun_out/ncu_bn.log 2>&1; tail -1 gpurun_out/ncu_bn.log
ls -la gpurun_out/*.ncu-rep
