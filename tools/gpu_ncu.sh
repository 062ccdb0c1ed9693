#!/bin/bash
mkdir -p gpurun_out
export ENDO_TC_DISABLE=8192   # single stream: per-kernel times are exclusive under the profiler anyway
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_final.csv python tools/profile_step.py 1 tf32x3 > gpurun_out/ncu_launch_final.log 2>&1; echo "launch list exit $?"
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dense_dgrad_tf32 -c 1 -o gpurun_out/prof_final_dgrad -f python tools/profile_step.py 1 tf32x3 > /dev/null 2>&1; echo "dgrad exit $?"
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dense_wgrad_bf16 -c 1 -o gpurun_out/prof_final_wgrad -f python tools/profile_step.py 1 tf32x3 > /dev/null 2>&1; echo "wgrad exit $?"
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dense_fwd_tf32 -s 54 -c 1 -o gpurun_out/prof_final_fwd -f python tools/profile_step.py 1 tf32x3 > /dev/null 2>&1; echo "fwd exit $?"
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:pw_gemm -c 1 -o gpurun_out/prof_final_pw -f python tools/profile_step.py 1 tf32x3 > /dev/null 2>&1; echo "pw exit $?"
unset ENDO_TC_DISABLE
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:warp_ -o gpurun_out/prof_final_warp -f python tools/profile_warp.py c5 1 > /dev/null 2>&1; echo "warp exit $?"
