#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu14.log 2>&1
grep -E "passed|failed|^E   +Assertion|FAILED" gpurun_out/pytest_gpu14.log | head
timeout 600 python bench.py --steps 5 --warmup 3 --math fp32 --no-cpu-baseline --no-e2e > gpurun_out/bench_fp32k.json 2> gpurun_out/bench_fp32k.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_fp32k.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
timeout 600 python bench.py --steps 5 --warmup 3 --math tf32 --no-cpu-baseline --no-e2e > gpurun_out/bench_tf32k.json 2> gpurun_out/bench_tf32k.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tf32k.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
