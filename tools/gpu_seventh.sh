#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_probe.py -m gpu -q --timeout=300 -s > gpurun_out/pytest_probe7.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_probe7.log
grep -E "passed|failed|FAILED|cycles per MMA|AssertionError" gpurun_out/pytest_probe7.log | head -40
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout=600 -s -k "tensor_core or tf32" > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu7.log
grep -E "passed|failed|^E   +Assertion|FAILED|tensor-core|tf32 forward" gpurun_out/pytest_gpu7.log | head -30
timeout 600 python bench.py --steps 5 --warmup 3 --math tf32 --no-cpu-baseline --no-e2e > gpurun_out/bench_tf32d.json 2> gpurun_out/bench_tf32d.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tf32d.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
