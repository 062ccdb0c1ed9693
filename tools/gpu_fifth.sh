#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/debug_net.py 2 128 160 fp32 > gpurun_out/debug_net_fp32b.log 2>&1; tail -12 gpurun_out/debug_net_fp32b.log
timeout 600 python tools/debug_net.py 2 128 160 tf32 > gpurun_out/debug_net_tf32b.log 2>&1; grep -E "^y |rel " gpurun_out/debug_net_tf32b.log | head -14; tail -2 gpurun_out/debug_net_tf32b.log
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_geometry.py -m gpu -q --timeout=600 > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu5.log
grep -E "passed|failed|^E   +Assertion|FAILED" gpurun_out/pytest_gpu5.log | head -30
timeout 600 python bench.py --steps 5 --warmup 3 --math tf32 --no-cpu-baseline --no-e2e > gpurun_out/bench_tf32b.json 2> gpurun_out/bench_tf32b.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tf32b.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
tail -3 gpurun_out/bench_tf32b.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_tf32.csv python tools/profile_step.py 1 tf32 > gpurun_out/ncu_launch_tf32.log 2>&1; echo "ncu launch list exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"dense_fwd_tf32|dense_dgrad_tf32" -s 2 -c 4 -o gpurun_out/prof_tc_r1 python tools/profile_step.py 1 tf32 > gpurun_out/ncu_tc.log 2>&1; echo "ncu tc exit $?"
