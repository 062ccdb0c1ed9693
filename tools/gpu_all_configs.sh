#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg_c2.json 2> gpurun_out/bench_cfg.err; echo "bench c2 exit $?"; tail -2 gpurun_out/bench_cfg.err
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_cfg_c5.json 2>> gpurun_out/bench_cfg.err; echo "bench c5 exit $?"; tail -2 gpurun_out/bench_cfg.err
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_cfg_c3.json 2>> gpurun_out/bench_cfg.err; echo "bench c3 exit $?"; tail -2 gpurun_out/bench_cfg.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_cfg_ref.json 2>> gpurun_out/bench_cfg.err; echo "ref exit $?"
python - <<'PY'
import json
for n in ("c2","c5","c3","ref"):
    try:
        d=json.load(open(f"gpurun_out/bench_cfg_{n}.json"))
        print(n, round(d["value"],2), round(d["ms_per_step"],2), d.get("e2e") and round(d["e2e"]["value"],2), d.get("clocks"))
        if n=="c2":
            print({k:v["ms_per_step"] for k,v in d["kernels"].items()})
            print({k:(round(v["value"],1), round(v["ms_per_step"],2)) for k,v in d["other_math_modes"].items()})
            print(d["warp_layer"]); print(d["cpu_baseline"])
    except Exception as e: print(n, "ERR", e)
PY
