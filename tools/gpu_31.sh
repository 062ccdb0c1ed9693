#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --math bf16x3 --no-cpu-baseline --no-extra --no-e2e --steps 10 --warmup 3 > gpurun_out/bench_31.json 2> gpurun_out/bench_31.err; echo "bench exit $?"
tail -8 gpurun_out/bench_31.err
