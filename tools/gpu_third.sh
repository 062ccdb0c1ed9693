#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_probe.py tests/test_gpu_net.py -m gpu -q --timeout=600 > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu3.log
grep -E "passed|failed|^E  |FAILED" gpurun_out/pytest_gpu3.log | head -40
# launch list of one profiled step (single pass, no replay)
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_fp32.csv python tools/profile_step.py 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launch list exit $?"
# full-set capture of the three heaviest conv kernels (one launch each at full resolution) and of the warp kernels
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_kernel|wgrad_kernel" -s 12 -c 9 -o gpurun_out/prof_conv_r1 python tools/profile_step.py 1 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"warp_fwd|warp_bwd|flow_fwd|flow_bwd" -c 8 -o gpurun_out/prof_geom_r1 python tools/profile_step.py 1 > gpurun_out/ncu_geom.log 2>&1; echo "ncu geom exit $?"
ls -la gpurun_out
