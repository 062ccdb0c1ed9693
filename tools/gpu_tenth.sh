#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout=600 -s -k "tensor_core or tf32" > gpurun_out/pytest_gpu10.log 2>&1
grep -E "passed|failed|^E   +Assertion|FAILED|tensor-core|tf32 forward" gpurun_out/pytest_gpu10.log | head
for d in 0 3; do ENDO_TC_DEBUG=$d timeout 300 python tools/time_fwd.py 5 2>&1 | tail -1; done
timeout 600 python bench.py --steps 5 --warmup 3 --math tf32 --no-cpu-baseline --no-e2e > gpurun_out/bench_tf32g.json 2> gpurun_out/bench_tf32g.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tf32g.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_tf32g.csv python tools/profile_step.py 1 tf32 > gpurun_out/ncu_launch_tf32g.log 2>&1; echo "ncu launch list exit $?"
