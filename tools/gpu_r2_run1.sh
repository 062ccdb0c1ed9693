#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_run1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r2_pytest1.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/r2_pytest1.log
grep -n "rel \|gradient\|run-to-run\|forward_pair" gpurun_out/r2_pytest1.log | head -80
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke1.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r2_smoke1.log
bash tools/gpu_sanitize.sh
ENDO_TC_DEBUG=4 timeout 200 python tools/trace_fwd.py tf32x3 > gpurun_out/r2_trace_fwd.log 2>&1; echo "trace exit $?"; cat gpurun_out/r2_trace_fwd.log | head -40
