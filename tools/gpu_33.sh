#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_x3_v5.csv python tools/profile_step.py 1 tf32x3 > gpurun_out/ncu_launch_v5.log 2>&1; echo "ncu launch list exit $?"
