"""clock64 trace of CTA (0,0) of the FIRST weight-gradient launch of a backward (denseBlocksUp.4.layers.3, Cin 180, 256x320 x16):
   ENDO_TC_DEBUG=8 ENDO_TC_DISABLE=8192 python tools/trace_wgrad.py"""
import ctypes, os, sys, torch   # needs a trace build: ENDO_BUILD_TRACE=1 python -m endo_b200.build --force
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ENDO_TC_DEBUG"] = "8"
os.environ["ENDO_TC_DISABLE"] = "8192"
import endo_b200
from endo_b200 import _lib
model = endo_b200.models.FCDenseNet57(1, math="tf32x3")
endo_b200.engine.kaiming_init_(model, seed=1)
model.cuda().train()
x = torch.rand(16, 3, 256, 320, device="cuda") * 2 - 1
for _ in range(2):
    y = model(x)
    y.sum().backward()
torch.cuda.synchronize()
# the LAST traced launch of the backward is the firstconv-side one; re-run stopping after the first wgrad is not possible from here,
# so trace the whole backward and read what the last weight-gradient launch (denseBlocksDown.0.layers.0, Cin 48) left
buf = (ctypes.c_longlong * 2048)()
_lib.check(_lib.lib().endo_debug_trace_read(buf, 2048), "trace")
t = list(buf)
n = int(t[0]); t0 = t[1]
print("tiles of CTA (0,0):", n)
print(" it | top->raw landed | ->planes free | ->act written | ->grad written+arrive | loop period || mma: ready->issued | tma issued rel top")
prev = None
for it in range(min(n, 40)):
    b = 16 + 8 * it
    if b + 7 >= 2048: break
    top, raw, free, act, end, mrdy, miss, tma = t[b:b + 8]
    per = (top - prev) if prev else 0
    prev = top
    print(f"{it:3d} | {raw-top:6d} | {free-raw:6d} | {act-free:6d} | {end-act:6d} | {per:6d} || {miss-mrdy:6d} (ready at +{mrdy-top}) | tma +{tma-top}")
