#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']); print('tc', d['tensor_core']['value'], d['tensor_core']['ms_per_step']); print('warp',d['warp_layer']); print('cpu',d['cpu_baseline']); print('clocks',d['clocks']); print('roof',d['roofline'])"
tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref2.json 2>> gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_ref2.json
