#!/bin/bash
# one ncu --set full capture: tools/gpu_ncu_one.sh <tag> <kernel regex> [skip]
mkdir -p gpurun_out
export ENDO_TC_DISABLE=8192
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$2 -s ${3:-0} -c 1 -o gpurun_out/$1 -f python tools/profile_step.py 1 tf32x3 > /dev/null 2>&1; echo "ncu exit $?"
