"""Time the tf32 network forward alone (CUDA events): python tools/time_fwd.py [iters]   (ENDO_TC_DEBUG experiments)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import endo_b200
from endo_b200 import _lib
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
model = endo_b200.models.FCDenseNet57(1, math="tf32")
endo_b200.engine.kaiming_init_(model, seed=1)
model.cuda().train()
x = torch.rand(16, 3, 256, 320, device="cuda") * 2 - 1
with torch.no_grad():
    for _ in range(2):
        model(x)
    torch.cuda.synchronize()
    with _lib.profile() as prof:
        for _ in range(iters):
            model(x)
print("ENDO_TC_DEBUG", os.environ.get("ENDO_TC_DEBUG", "0"), {k: round(v / iters, 3) for k, v in prof.ms.items() if v > 0})
