#!/bin/bash
mkdir -p gpurun_out
ENDO_TC_DISABLE=8192 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dense_wgrad_bf16 -c 1 -o gpurun_out/prof_wgrad_v7 -f python tools/profile_step.py 1 tf32x3 > gpurun_out/ncu_w7.log 2>&1; echo "wgrad exit $?"
