#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout=600 -s -k "tf32" > gpurun_out/pytest_gpu13.log 2>&1
grep -E "passed|failed|^E   +Assertion|FAILED|tensor-core|tf32 forward" gpurun_out/pytest_gpu13.log | head
ENDO_TC_DEBUG=4 timeout 300 python tools/trace_fwd.py > gpurun_out/trace_fwd3.log 2>&1; cat gpurun_out/trace_fwd3.log | tail -17
timeout 600 python bench.py --steps 5 --warmup 3 --math tf32 --no-cpu-baseline --no-e2e > gpurun_out/bench_tf32j.json 2> gpurun_out/bench_tf32j.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tf32j.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
