#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -s -k "tf32x3 or forward_backward_vs_oracle or growth16" > gpurun_out/pytest_27.log 2>&1; tail -3 gpurun_out/pytest_27.log
grep -n "^FAILED\|^E  \|rel err\|gradient error\|median\|max " gpurun_out/pytest_27.log | head -40
