#!/bin/bash
# round-2 final evidence run: GPU tests, smoke, full bench line, reference arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2_final_smoke.log
timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2_final_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_final_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'launches/step',d.get('gpu_launches_per_step'),'loss',d['config']['loss'])
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'modules',d['e2e']['modules_as_train_py']['value'])
for k,v in d['kernels'].items(): print('  %-18s %8.3f ms  %s'%(k,v['ms_per_step'],v.get('frac_of_hbm_peak')))
print('roofline',{k:d['roofline'][k] for k in ('kernel','frac','achieved','traffic')})
print('other_configs',{k:(v['value'],v['ms_per_step'],v['loss']) for k,v in d['other_configs'].items()})
print('other modes',{k:(v['value'],v['loss']) for k,v in d['other_math_modes'].items()})
print('warp',d['warp_layer']); print('cpu',d['cpu_baseline']); print('clocks',d['clocks'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; echo "ref exit $?"; cat gpurun_out/r2_final_ref.json | cut -c1-600
