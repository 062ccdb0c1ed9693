#!/bin/bash
# ncu --set full captures (with source counters) of the hot kernels of one bs8 256x320 step, math tf32x3, single stream.
mkdir -p gpurun_out
TAG=${1:-r2}
export ENDO_TC_DISABLE=8192
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:dense_dgrad_tf32 -c 1 -o gpurun_out/${TAG}_dgrad -f python tools/profile_step.py 1 tf32x3 > /dev/null 2>&1; echo "dgrad exit $?"
timeout 400 $NCU -k regex:dense_wgrad_bf16 -c 1 -o gpurun_out/${TAG}_wgrad -f python tools/profile_step.py 1 tf32x3 > /dev/null 2>&1; echo "wgrad exit $?"
timeout 400 $NCU -k regex:dense_fwd_tf32 -s 58 -c 1 -o gpurun_out/${TAG}_fwd180 -f python tools/profile_step.py 1 tf32x3 > /dev/null 2>&1; echo "fwd exit $?"
ls -la gpurun_out/${TAG}_*.ncu-rep
