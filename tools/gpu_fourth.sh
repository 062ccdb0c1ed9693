#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/debug_net.py 2 128 160 tf32 > gpurun_out/debug_net_tf32.log 2>&1; echo "debug exit $?" >> gpurun_out/debug_net_tf32.log
head -30 gpurun_out/debug_net_tf32.log; tail -4 gpurun_out/debug_net_tf32.log
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout=600 -s > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu4.log
grep -E "passed|failed|^E  |FAILED|tf32 forward" gpurun_out/pytest_gpu4.log | head -30
timeout 600 python bench.py --steps 5 --warmup 3 --math tf32 --no-cpu-baseline --no-e2e > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tf32.json')); print(d['value'], d['ms_per_step']); print(json.dumps(d['kernels'], indent=0))"
tail -3 gpurun_out/bench_tf32.err
