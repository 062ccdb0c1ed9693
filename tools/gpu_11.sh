#!/bin/bash
mkdir -p gpurun_out
ENDO_TC_DEBUG=4 timeout 300 python tools/trace_fwd.py > gpurun_out/trace_fwd.log 2>&1; cat gpurun_out/trace_fwd.log | tail -20
timeout 600 python bench.py --steps 5 --warmup 3 --math tf32 --no-cpu-baseline --no-e2e > gpurun_out/bench_tf32h.json 2> gpurun_out/bench_tf32h.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tf32h.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
timeout 600 python bench.py --steps 5 --warmup 3 --math fp32 --no-cpu-baseline --no-e2e > gpurun_out/bench_fp32h.json 2> gpurun_out/bench_fp32h.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_fp32h.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout=600 > gpurun_out/pytest_gpu11.log 2>&1
grep -E "passed|failed|^E   +Assertion|FAILED" gpurun_out/pytest_gpu11.log | head
