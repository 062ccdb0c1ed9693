#!/bin/bash
# 2-GPU run of the bench exactly as the driver launches it (torchrun, one rank per GPU, NCCL)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "n2 exit $?"; tail -5 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n2.json') if l.startswith('{')][-1])
print('n_gpus',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'modules',d['e2e']['modules_as_train_py']['value'],'loss',d['config']['loss'])
print('other_configs',{k:(v['value'],v['ms_per_step']) for k,v in d['other_configs'].items()})
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-300
