#!/bin/bash
# ncu launch list (gpu__time_duration) of ONE optimisation step, single stream: tools/gpu_launch_list.sh <tag>
mkdir -p gpurun_out
export ENDO_TC_DISABLE=8192
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/$1_launches.csv python tools/profile_step.py 1 tf32x3 > gpurun_out/$1_launch.log 2>&1; echo "launch list exit $?"
