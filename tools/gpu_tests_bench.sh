#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_full.log 2>&1; tail -3 gpurun_out/pytest_full.log
grep -n "^FAILED\|^E  " gpurun_out/pytest_full.log | head -20
timeout 600 python bench.py --no-cpu-baseline --no-extra --no-e2e --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"
tail -3 gpurun_out/bench_full.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()}); print(d['roofline'])"
