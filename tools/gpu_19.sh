#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_19.log 2>&1; tail -3 gpurun_out/pytest_19.log
grep -n "^FAILED\|^E  " gpurun_out/pytest_19.log | head -20
timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_19.json 2> gpurun_out/bench_19.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_19.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()}); print('tc', d['tensor_core']['value'], d['tensor_core']['loss'], d['config']['loss']); print(d['warp_layer'])"
