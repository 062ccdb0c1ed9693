#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_tc_probe.py -m gpu -q -s -k "tf32 or tensor_core" > gpurun_out/pytest_26.log 2>&1; tail -3 gpurun_out/pytest_26.log
grep -n "^FAILED\|^E  \|rel err\|gradient error\|median\|max " gpurun_out/pytest_26.log | head -40
