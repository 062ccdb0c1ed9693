"""DepthWarpingLayer forward + backward alone at a bench size (for ncu):  python tools/profile_warp.py [c2|c5] [iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import endo_b200  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
b, h, w = {"c2": (8, 256, 320), "c5": (16, 512, 640)}[cfg]
dev = torch.device("cuda", 0)
batch = endo_b200.synthetic.make_batch(b, h, w, seed=7)
d1, d2 = endo_b200.synthetic.jitter_depths(batch, seed=8)
args = [batch[k].to(dev) for k in ("boundaries", "translations_1_wrt_2", "rotations_1_wrt_2", "intrinsics")]
x1, x2 = d1.to(dev).requires_grad_(True), d2.to(dev).requires_grad_(True)
gw = torch.randn(b, 1, h, w, device=dev)
layer = endo_b200.models.DepthWarpingLayer(epsilon=1.0e-8)
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
wd, _ = layer([x1, x2] + args)
torch.autograd.grad(wd, [x1, x2], gw)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(iters):
    flush.zero_()
    wd, _ = layer([x1, x2] + args)
    flush.zero_()
    torch.autograd.grad(wd, [x1, x2], gw)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
