"""Per-phase clock64() trace of CTA 0 of ONE tensor-core DenseLayer forward launch (development aid).
   ENDO_TC_DEBUG=4 python tools/trace_fwd.py [tf32|tf32x3]"""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ENDO_TC_DEBUG"] = os.environ.get("ENDO_TC_DEBUG", "4")
import endo_b200
from endo_b200 import _lib
model = endo_b200.models.FCDenseNet57(1, math=(sys.argv[1] if len(sys.argv) > 1 else "tf32"))
endo_b200.engine.kaiming_init_(model, seed=1)
model.cuda().train()
x = torch.rand(16, 3, 256, 320, device="cuda") * 2 - 1
with torch.no_grad():
    model(x); model(x)
buf = (ctypes.c_longlong * 512)()
_lib.check(_lib.lib().endo_debug_trace_read(buf, 512), "trace")
t = list(buf)
n = int(t[0]); t0 = t[1]
print("last traced launch = finalmost DenseLayer (up-block, level 0); chunks:", n)
print(f"setup->first chunk: {t[16] - t0}; epilogue wait {t[3] - t[2]}; epilogue {t[4] - t[3]} "
      f"(edge pass {t[6] - t[3]}, barrier {t[7] - t[6]}, combine+store {t[8] - t[7]}, statistics {t[4] - t[8]}); total {t[5] - t0} cycles")
print("combine+store of warp 0, per M-block [TMEM loads done, stores issued] rel. to its start:", [t[i] - t[7] for i in range(9, 15)])
print("chunk: prod[wait_empty, loads+stores, weights, fence+arrive] | mma[wait_full, issue] | prod start rel, mma start rel")
for c in range(n):
    b = 16 + c * 8
    print(f"{c:2d}: {t[b+1]-t[b]:6d} {t[b+2]-t[b+1]:6d} {t[b+3]-t[b+2]:6d} {t[b+4]-t[b+3]:6d} | {t[b+6]-t[b+5]:6d} {t[b+7]-t[b+6]:6d} | {t[b]-t0:7d} {t[b+5]-t0:7d}")
