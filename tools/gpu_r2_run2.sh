#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_c2.py -q -s -x > gpurun_out/r2_parity_c2.log 2>&1; echo "parity exit $?"; tail -3 gpurun_out/r2_parity_c2.log
grep -n "rel \|gradient\|run-to-run\|forward_pair\|config 3\|eager\|Error\|error" gpurun_out/r2_parity_c2.log | cut -c1-400 | head -70
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench exit $?"; tail -5 gpurun_out/r2_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench2.json'))
print('value',d['value'],'ms',d['ms_per_step'],'launches/step',d.get('gpu_launches_per_step'),'loss',d['config']['loss'])
print('e2e',d['e2e'])
print({k:(v['ms_per_step'],v.get('frac_of_hbm_peak')) for k,v in d['kernels'].items()})
print('roofline',{k:d['roofline'][k] for k in ('kernel','frac','achieved')})
print('other_configs',d['other_configs'])
print('other modes',{k:(v['value'],v['loss']) for k,v in d['other_math_modes'].items()})
print('warp',d['warp_layer']); print('cpu',d['cpu_baseline']); print('clocks',d['clocks'])
PY
timeout 900 python bench.py --impl reference-gpu --steps 10 --warmup 3 > gpurun_out/r2_refgpu.json 2> gpurun_out/r2_refgpu.err; echo "refgpu exit $?"; tail -3 gpurun_out/r2_refgpu.err; cat gpurun_out/r2_refgpu.json
