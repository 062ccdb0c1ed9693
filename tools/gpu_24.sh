#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_tf32_v4.csv python tools/profile_step.py 1 tf32 > gpurun_out/ncu_launch_v4.log 2>&1; echo "ncu launch list exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:warp_ -o gpurun_out/prof_warp_c5 -f python tools/profile_warp.py c5 1 > gpurun_out/ncu_warp.log 2>&1; echo "ncu warp exit $?"
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
tail -3 gpurun_out/bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','e2e','clocks','gpu_launches')})"
