#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline --no-extra --no-e2e --steps 10 --warmup 3 > gpurun_out/bench_40.json 2> gpurun_out/bench_40.err; echo "bench exit $?"
tail -3 gpurun_out/bench_40.err
python -c "
import json; d=json.load(open('gpurun_out/bench_40.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
