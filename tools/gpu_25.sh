#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -s -k "tf32x3" > gpurun_out/pytest_25.log 2>&1; tail -3 gpurun_out/pytest_25.log
grep -n "^FAILED\|^E  \|tf32x3" gpurun_out/pytest_25.log | head -20
timeout 600 python bench.py --math tf32x3 --no-cpu-baseline --no-extra --no-e2e --steps 10 --warmup 3 > gpurun_out/bench_25.json 2> gpurun_out/bench_25.err; echo "bench exit $?"
tail -3 gpurun_out/bench_25.err
python -c "
import json; d=json.load(open('gpurun_out/bench_25.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()}); print(d['config']['loss'])"
