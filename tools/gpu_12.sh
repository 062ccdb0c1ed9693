#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout=600 -s -k "tensor_core or tf32" > gpurun_out/pytest_gpu12.log 2>&1
grep -E "passed|failed|^E   +Assertion|FAILED|tensor-core|tf32 forward" gpurun_out/pytest_gpu12.log | head
ENDO_TC_DEBUG=4 timeout 300 python tools/trace_fwd.py > gpurun_out/trace_fwd2.log 2>&1; cat gpurun_out/trace_fwd2.log | tail -16
timeout 600 python bench.py --steps 5 --warmup 3 --math tf32 --no-cpu-baseline --no-e2e > gpurun_out/bench_tf32i.json 2> gpurun_out/bench_tf32i.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tf32i.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_tf32i.csv python tools/profile_step.py 1 tf32 > gpurun_out/ncu_launch_tf32i.log 2>&1; echo "ncu launch list exit $?"
