#!/usr/bin/env python
"""bench.py -- headline benchmark of the endo-depth-b200 hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c5]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one optimisation step of /root/reference/train.py:272-328 on one synthetic batch:
FCDenseNet57 forward on both images of every pair, DepthScalingLayer x2, FlowfromDepthLayer x2 +
SparseMaskedL1Loss x2, DepthWarpingLayer x2 + NormalizedDistanceLoss x2 (dcl_weight 5, sfl_weight 20),
backward, clip_grad_norm_(10) and SGD(momentum 0.9); with N > 1 one gradient all-reduce per step.
Metric: image-pairs/sec (whole job).  Default workload = BASELINE.json configs[1]: bs8/GPU, 256x320, fp32.

Prints ONE JSON line (rank 0):
  value     device-resident throughput: inputs already in HBM, fused pair forward + fused optimiser tail,
            timed with CUDA events, barrier + synchronize on both sides, max over ranks
  e2e       the same step through the reference-facing modules exactly as train.py drives them
            (net(colors_1); net(colors_2); torch.optim.SGD; clip_grad_norm_; loss.item()), host buffers in
            pinned memory, 16 H2D copies + 1 D2H read per step inside the timed region
  roofline  dominant kernel class (per-category CUDA-event timing inside this process, endo_prof_*)
  cpu_baseline  the oracle port (CPU restatement of the reference path) on the host cores, bounded sample
`--impl reference` times that CPU path alone (the reference is pure PyTorch-on-CPU for this tier).
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (batch per GPU, H, W, math, description)
    "c2": (8, 256, 320, "fp32", "1xB200 bs8 256x320 synthetic pairs, FCDenseNet57, full loss stack (dcl 5, sfl 20), fp32"),
    "c5": (16, 512, 640, "fp32", "bs16/GPU 512x640 (downsampling 2.0), warp-gather stress"),
}
METRIC = "image-pairs/sec fwd+bwd @256x320 bs8"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def conv_flops_per_image(h, w):
    """2*Cin*Cout*k*k*Hout*Wout over the 56 convs of FCDenseNet57, split by kernel class."""
    g, first = 12, 48
    dense, trans = 0.0, 0.0
    trans += 2.0 * 3 * first * 9 * h * w                        # firstconv
    cur, res, skips = first, 1, []
    for _ in range(5):
        for j in range(4):
            dense += 2.0 * (cur + j * g) * g * 9 * (h // res) * (w // res)
        cur += 4 * g
        skips.append((cur, res))
        trans += 2.0 * cur * cur * (h // res) * (w // res)      # TransitionDown 1x1
        res *= 2
    for j in range(4):                                           # bottleneck
        dense += 2.0 * (cur + j * g) * g * 9 * (h // res) * (w // res)
    for _ in range(5):
        cs, res = skips.pop()
        trans += 2.0 * 48 * 48 * 9 * (h // res) * (w // res)    # TransitionUp 3x3 at the upsampled size
        for j in range(4):
            dense += 2.0 * (48 + cs + j * g) * g * 9 * (h // res) * (w // res)
    final = 2.0 * 192 * h * w
    return dense, trans, final


def pick_cpu_threads(h, w):
    """The CPU arm should use as many host threads as actually help: torch's default on a 128-logical-CPU
    box oversubscribes these small tensors badly.  Time one bs1 forward at a few pool sizes and keep the best."""
    from oracle import net as onet
    import endo_b200
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    cands = sorted({c for c in (4, 8, 16, 32, 64, avail) if c <= avail})
    state = onet.init_state(onet.FCDENSENET57, seed=1)
    x = endo_b200.synthetic.make_batch(1, h, w, seed=1)["colors_1"]
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            onet.forward(state, x, onet.FCDENSENET57, True, {})
            t0 = time.perf_counter()
            onet.forward(state, x, onet.FCDENSENET57, True, {})
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best, avail


def run_reference(args, rank, world):
    """CPU arm: the oracle port of train.py:272-325 (reference semantics, torch CPU ops) on all host threads."""
    if rank != 0:
        return
    from oracle import net as onet, step as ostep, geometry as ogeo
    import endo_b200
    onet.LIBRARY_OPS = ogeo.LIBRARY_OPS = True       # the torch library ops the reference itself calls
    _, h, w, _, desc = CONFIGS[args.config]
    cores, avail = pick_cpu_threads(h, w)
    sample_b = 1
    state = onet.init_state(onet.FCDENSENET57, seed=10085)
    batch = endo_b200.synthetic.make_batch(sample_b, h, w, seed=10085)
    mom = {}
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        loss, dcl, sfl, grads, new_buf, _ = ostep.forward_backward(state, batch, onet.FCDENSENET57, 5.0, 20.0)
        ostep.clip_and_sgd(state, grads, mom, lr=1e-3)
        state.update(new_buf)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample_b * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": desc, "sample": f"bs{sample_b} {h}x{w} per step (same per-pair work)"},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                             "sample": f"{len(times)} steps of bs{sample_b} {h}x{w}: oracle port of the reference "
                                       f"path (torch CPU ops, best of several pool sizes = {cores} threads, "
                                       f"{avail} logical CPUs available)"},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(h, w, budget_s=20.0):
    from oracle import net as onet, step as ostep, geometry as ogeo
    import endo_b200
    onet.LIBRARY_OPS = ogeo.LIBRARY_OPS = True       # the torch library ops the reference itself calls
    cores, avail = pick_cpu_threads(h, w)
    state = onet.init_state(onet.FCDENSENET57, seed=10085)
    batch = endo_b200.synthetic.make_batch(1, h, w, seed=10085)
    ostep.forward_backward(state, batch, onet.FCDENSENET57, 5.0, 20.0)      # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < 5 and (time.perf_counter() - t_start) < budget_s:
        t0 = time.perf_counter()
        ostep.forward_backward(state, batch, onet.FCDENSENET57, 5.0, 20.0)
        times.append(time.perf_counter() - t0)
    med = sorted(times)[len(times) // 2]
    return {"value": 1.0 / med, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"median of {len(times)} fwd+bwd steps of bs1 {h}x{w} (BASELINE config 1) on the oracle port, "
                      f"{cores} torch threads (best of several pool sizes; {avail} logical CPUs available)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--math", default=None, choices=["fp32", "tf32", "bf16"],
                    help="override the arithmetic of the conv path (default: the config's, fp32)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end arm (kernel development runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import endo_b200
    from endo_b200 import _lib, ddp, train_step
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: endo_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        ddp.init_from_env("nccl")
    bsz, h, w, math_mode, desc = CONFIGS[args.config]
    if args.math:
        math_mode = args.math
    peaks = measured_peaks()

    torch.manual_seed(10085 + rank)
    model = endo_b200.models.FCDenseNet57(n_classes=1, math=math_mode)
    endo_b200.engine.kaiming_init_(model, seed=10085)           # identical weights on every rank
    model.to(dev).train()
    host = endo_b200.synthetic.make_batch(bsz, h, w, seed=10085 + rank)
    keys = endo_b200.synthetic.BATCH_KEYS_H2D
    host = {k: host[k].pin_memory() for k in keys}
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------ device-resident arm (value)
    pg = dist.group.WORLD if world > 1 else None
    fused = train_step.TrainStep(model, h, w, lr=1e-4, momentum=0.9, max_norm=10.0, dcl_weight=5.0, sfl_weight=20.0,
                                 pair=True, process_group=pg)
    for _ in range(args.warmup):
        fused.step(resident)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, _, _ = fused.step(resident)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - launches0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    final_loss = float(loss)
    ms_per_step = ms_total / args.steps
    value = world * bsz * args.steps / (ms_total / 1e3)

    # ------------------------------------------------------------------ end-to-end arm (e2e)
    model2 = endo_b200.models.FCDenseNet57(n_classes=1, math=math_mode)
    endo_b200.engine.kaiming_init_(model2, seed=10085)
    model2.to(dev).train()
    stack = train_step.LossStack(h, w, dcl_weight=5.0, sfl_weight=20.0)
    opt = torch.optim.SGD(model2.parameters(), lr=1e-4, momentum=0.9)          # train.py:202

    def e2e_step():
        cb = {k: host[k].to(dev, non_blocking=True) for k in keys}             # train.py:254-270
        lv, _, _, _ = stack.loss(model2, cb)                                   # :272-315 (two separate net() calls)
        val = lv.item()                                                        # :317 device->host read
        if train_step.is_bad(val):
            opt.zero_grad()
            return val
        opt.zero_grad()
        lv.backward()
        if world > 1:
            ddp.allreduce_gradients(model2)
        torch.nn.utils.clip_grad_norm_(model2.parameters(), 10.0)              # :327
        opt.step()                                                             # :328
        return val

    e2e_steps = 0 if args.no_e2e else args.steps
    for _ in range(args.warmup if e2e_steps else 0):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(max(e2e_steps, 1)):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) * (args.steps / max(e2e_steps, 1))
    e2e_value = world * bsz * args.steps / (e2e_ms / 1e3)

    # ------------------------------------------------------------------ per-kernel-class timing (roofline)
    prof_steps = 3
    with _lib.profile() as prof:
        for _ in range(prof_steps):
            fused.step(resident)
    barrier()
    dense_f, trans_f, final_f = conv_flops_per_image(h, w)
    imgs = 2 * bsz                                                             # both images of every pair
    P = bsz * h * w
    cat_ms = {k: v / prof_steps for k, v in prof.ms.items()}
    cat_n = {k: v // prof_steps for k, v in prof.counts.items()}
    # algorithmic work per step for each class (fwd FLOPs of the convs; dgrad and wgrad redo the same MACs)
    work_flops = {"conv_dense_fwd": dense_f * imgs, "conv_trans_fwd": trans_f * imgs,
                  "conv_dgrad": (dense_f + trans_f - 2.0 * 3 * 48 * 9 * h * w) * imgs,
                  "conv_wgrad": (dense_f + trans_f) * imgs}
    work_bytes = {"depth_warp": 2 * 44.0 * P, "flow_from_depth": 2 * 36.0 * P, "depth_scale": 2 * 48.0 * P}
    dominant = max(("conv_dense_fwd", "conv_dgrad", "conv_wgrad", "conv_trans_fwd"), key=lambda k: cat_ms.get(k, 0.0))
    dom_ms = cat_ms[dominant]
    dom_launches = max(cat_n[dominant], 1)
    achieved_tf = work_flops[dominant] / (dom_ms * 1e-3) / 1e12
    peak_tf = peaks["bf16_tflops_sustained"]
    roofline = {"kernel": dominant, "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": None, "peak_source": peaks["source"] + " (bf16 sustained)",
                "launches_per_step": dom_launches, "avg_launch_ms": dom_ms / dom_launches,
                "flops_per_launch": work_flops[dominant] / dom_launches,
                "note": "fp32 FFMA implicit-GEMM path (no tensor pipe yet); fraction is of the measured bf16 tensor peak; "
                        "fp32 FFMA nominal peak on B200 is ~75 TFLOP/s",
                "share_of_step": dom_ms / max(sum(cat_ms.values()), 1e-9)}
    warp_ms = cat_ms.get("depth_warp", 0.0)
    kernels = {k: {"ms_per_step": round(v, 4), "launch_sites": cat_n[k]} for k, v in cat_ms.items()}
    for k, byts in work_bytes.items():
        if cat_ms.get(k, 0.0) > 0:
            gbs = byts / (cat_ms[k] * 1e-3) / 1e9
            kernels[k].update({"algorithmic_GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm_gbs"], 4)})
    for k, fl in work_flops.items():
        if cat_ms.get(k, 0.0) > 0:
            kernels[k].update({"TFLOPps": round(fl / (cat_ms[k] * 1e-3) / 1e12, 2)})

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(h, w)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": math_mode, "data": "synthetic",
                "config": {"workload": desc, "batch_per_gpu": bsz, "global_batch": bsz * world, "height": h, "width": w,
                           "model": "FCDenseNet57 (random Kaiming init)", "parallelism": f"dp{world}",
                           "l2": "per-step working set (~3 GB of activations and gradients) >> 126 MB L2: no explicit flush",
                           "loss": final_loss},
                "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_ms / args.steps,
                        "path": "reference-facing nn.Modules as train.py:254-328 drives them, pinned host buffers"},
                "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline,
                "kernels": kernels, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
