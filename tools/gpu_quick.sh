#!/bin/bash
# quick kernel-iteration loop: GPU tests (fail fast) + short bench with the per-class breakdown
mkdir -p gpurun_out
TAG=${1:-quick}
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log; grep -n "^FAILED\|^E  " gpurun_out/${TAG}_pytest.log | head -20
timeout 600 python bench.py --no-cpu-baseline --no-extra --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'launches/step',d.get('gpu_launches_per_step'),'loss',d['config']['loss'])
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'modules',d['e2e']['modules_as_train_py']['value'])
for k,v in d['kernels'].items(): print('  %-18s %8.3f ms  %s'%(k,v['ms_per_step'],v.get('frac_of_hbm_peak')))
print('sum', sum(v['ms_per_step'] for k,v in d['kernels'].items() if 'combined' not in k))
PY
