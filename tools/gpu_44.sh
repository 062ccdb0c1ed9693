#!/bin/bash
mkdir -p gpurun_out
ENDO_TC_DEBUG=4 timeout 300 python tools/trace_fwd.py tf32x3 2>&1 | tail -27 | head -3
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q -k "x3 or full_train_step or forward_backward or splitk or tf32_tensor" 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --no-extra --no-e2e --steps 10 --warmup 3 > gpurun_out/bench_44.json 2> gpurun_out/bench_44.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_44.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_step'] for k,v in d['kernels'].items()})"
