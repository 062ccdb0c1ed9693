#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_tc_probe.py -m gpu -q --timeout=600 -s > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu8.log
grep -E "passed|failed|^E   +Assertion|FAILED|tensor-core|tf32 forward" gpurun_out/pytest_gpu8.log | head -30
timeout 600 python bench.py --steps 5 --warmup 3 --math tf32 --no-cpu-baseline --no-e2e > gpurun_out/bench_tf32e.json 2> gpurun_out/bench_tf32e.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tf32e.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_tf32e.csv python tools/profile_step.py 1 tf32 > gpurun_out/ncu_launch_tf32e.log 2>&1; echo "ncu launch list exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"dense_fwd_tf32|dense_dgrad_tf32|dense_wgrad_bf16" -s 2 -c 3 -o gpurun_out/prof_tc_r1e python tools/profile_step.py 1 tf32 > gpurun_out/ncu_tc_e.log 2>&1; echo "ncu tc exit $?"
