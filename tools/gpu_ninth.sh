#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_probe.py -m gpu -q --timeout=300 -s -k "cost" > gpurun_out/pytest_probe9.log 2>&1
grep -E "passed|failed|FAILED|cycles per MMA|Error" gpurun_out/pytest_probe9.log | head -40
for d in 0 1 2 3; do ENDO_TC_DEBUG=$d timeout 300 python tools/time_fwd.py 5 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --timeout=600 -s -k "tensor_core or tf32" > gpurun_out/pytest_gpu9.log 2>&1
grep -E "passed|failed|^E   +Assertion|FAILED|tensor-core|tf32 forward" gpurun_out/pytest_gpu9.log | head
timeout 600 python bench.py --steps 5 --warmup 3 --math tf32 --no-cpu-baseline --no-e2e > gpurun_out/bench_tf32f.json 2> gpurun_out/bench_tf32f.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tf32f.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
