mkdir -p gpurun_out
ENDO_TC_DISABLE=8192 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_final.csv python tools/profile_step.py 1 tf32x3 > gpurun_out/ncu_launch_final.log 2>&1; echo "launch list exit $?"
