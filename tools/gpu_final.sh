#!/bin/bash
# round-end style check: GPU parity suite, smoke(), default bench line, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; tail -2 gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
( time timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | grep real; echo "bench exit $?"; tail -2 gpurun_out/bench_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print({k:d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','dtype','gpu_launches','clocks')}); print(d['e2e']); print({k:d['roofline'][k] for k in ('kernel','bound','achieved','peak','frac','traffic')}); print(d['cpu_baseline'])"
