#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity_c2.py -q -x -s -k "test_network_forward_bs8_256x320_pair_and_separate and tf32x3" > gpurun_out/fwd2_first.log 2>&1; echo "first exit $?"; tail -12 gpurun_out/fwd2_first.log | cut -c1-300
