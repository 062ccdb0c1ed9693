#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dense_wgrad_bf16 -c 1 -o gpurun_out/prof_wgrad_v5 -f python tools/profile_step.py 1 tf32x3 > gpurun_out/ncu_w.log 2>&1; echo "wgrad exit $?"
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dense_dgrad_tf32 -c 1 -o gpurun_out/prof_dgrad_v5 -f python tools/profile_step.py 1 tf32x3 > gpurun_out/ncu_d.log 2>&1; echo "dgrad exit $?"
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dense_fwd_tf32 -s 3 -c 1 -o gpurun_out/prof_fwd_v5 -f python tools/profile_step.py 1 tf32x3 > gpurun_out/ncu_f.log 2>&1; echo "fwd exit $?"
