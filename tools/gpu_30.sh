#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -s -k "x3" > gpurun_out/pytest_30.log 2>&1; tail -3 gpurun_out/pytest_30.log
grep -n "^FAILED\|^E  \|rel err\|gradient error\|forward vs" gpurun_out/pytest_30.log | cut -c1-250 | head -40
for m in bf16x3 tf32x3; do
timeout 600 python bench.py --math $m --no-cpu-baseline --no-extra --no-e2e --steps 10 --warmup 3 > gpurun_out/bench_30_$m.json 2> gpurun_out/bench_30.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_30_$m.json')); print('$m', d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
done
