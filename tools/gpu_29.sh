#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_probe.py -m gpu -q -s -k "mma_cost" 2>&1 | grep -i "cycles per\|passed\|failed"
ENDO_TC_DEBUG=4 timeout 300 python tools/trace_fwd.py tf32x3 2>&1 | tail -30
timeout 300 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "tf32x3" 2>&1 | tail -3
