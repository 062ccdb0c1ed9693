#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --math tf32 --no-e2e --no-extra --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_tf32_16.json 2> gpurun_out/bench_tf32_16.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tf32_16.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
