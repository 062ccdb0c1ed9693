#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_23.log 2>&1; tail -3 gpurun_out/pytest_23.log
grep -n "^FAILED\|^E  " gpurun_out/pytest_23.log | head -20
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_23.json 2> gpurun_out/bench_23.err; echo "bench exit $?"
tail -3 gpurun_out/bench_23.err
python -c "
import json; d=json.load(open('gpurun_out/bench_23.json')); print(d['value'], d['ms_per_step'], d['e2e']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()}); print(d['roofline']); tc=d['tensor_core']; print(tc['value'], tc['ms_per_step'], {k:v['ms_per_step'] for k,v in tc['kernels'].items()}); print(d['warp_layer']); print(d['cpu_baseline'])"
