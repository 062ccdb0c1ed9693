#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_geometry.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python tools/time_warp.py
