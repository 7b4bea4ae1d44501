#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python scripts/profile_step.py > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
wc -l gpurun_out/launches_r01.csv
