#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -s -k "tensor_core_backward or full_train_step or forward_backward_vs_oracle" > gpurun_out/pytest_41.log 2>&1; tail -3 gpurun_out/pytest_41.log
grep -n "^FAILED\|^E  \|tensor-core" gpurun_out/pytest_41.log | cut -c1-200 | head -20
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_x3_v6.csv python tools/profile_step.py 1 tf32x3 > gpurun_out/ncu_launch_v6.log 2>&1; echo "ncu launch list exit $?"
