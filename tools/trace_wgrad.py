"""Per-phase clock64() trace of CTA (0,0) of the LAST tensor-core weight-gradient launch of one backward (development aid):
   ENDO_TC_DEBUG=8 python tools/trace_wgrad.py   -> the last dense wgrad launch = denseBlocksDown.0.layers.0 (Cin 48, level 0)"""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ENDO_TC_DEBUG"] = os.environ.get("ENDO_TC_DEBUG", "8")
os.environ["ENDO_TC_DISABLE"] = "8192"
import endo_b200
from endo_b200 import _lib
model = endo_b200.models.FCDenseNet57(1, math="tf32")
endo_b200.engine.kaiming_init_(model, seed=1)
model.cuda().train()
x = torch.rand(16, 3, 256, 320, device="cuda") * 2 - 1
for _ in range(2):
    y = model(x)
    y.sum().backward()
buf = (ctypes.c_longlong * 2048)()
_lib.check(_lib.lib().endo_debug_trace_read(buf, 2048), "trace")
t = list(buf)
n = int(t[0])
print("tiles of CTA 0:", n)
print("tile: prod[wait_empty, A staging, G staging] | mma[wait_full, issue->commit] | prod start, mma start (rel. to first)")
t0 = t[16]
for it in range(min(n, 24)):
    b = 16 + it * 8
    print(f"{it:2d}: {t[b+1]-t[b]:6d} {t[b+2]-t[b+1]:6d} {t[b+3]-t[b+2]:6d} | {t[b+5]-t[b+4]:6d} {t[b+6]-t[b+5]:6d} | {t[b]-t0:7d} {t[b+4]-t0:7d}")
