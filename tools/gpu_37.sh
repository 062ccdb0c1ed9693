#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --no-extra --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_37.json 2> gpurun_out/bench_37.err; echo "bench exit $?"
tail -3 gpurun_out/bench_37.err
python -c "
import json; d=json.load(open('gpurun_out/bench_37.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])"
