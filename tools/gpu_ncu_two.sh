#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2}
export ENDO_TC_DISABLE=8192
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:dense_dgrad_tf32 -c 1 -o gpurun_out/${TAG}_dgrad -f python tools/profile_step.py 1 tf32x3 > /dev/null 2>&1; echo "dgrad exit $?"
timeout 400 $NCU -k regex:dense_wgrad_bf16 -c 1 -o gpurun_out/${TAG}_wgrad -f python tools/profile_step.py 1 tf32x3 > /dev/null 2>&1; echo "wgrad exit $?"
