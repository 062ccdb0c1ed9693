#!/bin/bash
# first GPU bring-up: stage-by-stage net debug, then the GPU test-suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python tools/debug_net.py > gpurun_out/debug_net.log 2>&1; echo "debug exit $?" >> gpurun_out/debug_net.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/debug_net.log; tail -15 gpurun_out/pytest_gpu.log
