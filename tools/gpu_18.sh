#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_18.log 2>&1; tail -5 gpurun_out/pytest_18.log
grep -n "tensor-core\|TransitionDown tc\|3xTF32\|tf32 MMA vs\|AssertionError\|^FAILED\|^E  " gpurun_out/pytest_18.log | head -40
