#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:warp_ -o gpurun_out/prof_warp_c5_v2 -f python tools/profile_warp.py c5 1 > gpurun_out/ncu_warp2.log 2>&1; echo "ncu warp exit $?"
