#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -s -k "x3 or full_train_step" > gpurun_out/pytest_39.log 2>&1; tail -3 gpurun_out/pytest_39.log
grep -n "^FAILED\|^E  \|rel err\|gradient error\|forward vs" gpurun_out/pytest_39.log | cut -c1-220 | head -20
timeout 600 python bench.py --no-cpu-baseline --no-extra --no-e2e --steps 10 --warmup 3 > gpurun_out/bench_39.json 2> gpurun_out/bench_39.err; echo "bench exit $?"
tail -3 gpurun_out/bench_39.err
python -c "
import json; d=json.load(open('gpurun_out/bench_39.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
