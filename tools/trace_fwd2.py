"""clock64 trace of CTA 0 of the LAST persistent forward launch (denseBlocksUp.4.layers.3, Cin 180, 16 x 256x320):
   python tools/trace_fwd2.py"""
import ctypes, os, sys, torch   # needs a trace build: ENDO_BUILD_TRACE=1 python -m endo_b200.build --force
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ENDO_TC_DEBUG"] = "16"
import endo_b200
from endo_b200 import _lib
model = endo_b200.models.FCDenseNet57(1, math="tf32x3")
endo_b200.engine.kaiming_init_(model, seed=1)
model.cuda().train()
x = torch.rand(16, 3, 256, 320, device="cuda") * 2 - 1
with torch.no_grad():
    model(x); model(x)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 2048)()
_lib.check(_lib.lib().endo_debug_trace_read(buf, 2048), "trace")
t = list(buf)
nch, t0, ntl = int(t[0]), t[1], int(t[2])
print(f"chunks per tile {nch}, tiles of CTA 0: {ntl}")
print("  j | top->raw landed | ->stage free | ->planes written | period || mma: ready rel top, issue time | tma issued rel top")
prev = None
for j in range(min(3 * nch, 70)):
    b = 16 + 8 * j
    top, raw, free, done, mrdy, miss, tma = t[b:b + 7]
    per = top - prev if prev else 0
    prev = top
    tag = " <- tile start" if j % nch == 0 else ""
    print(f"{j:3d} | {raw-top:6d} | {free-raw:6d} | {done-free:6d} | {per:6d} || +{mrdy-top:6d} {miss-mrdy:6d} | {tma-top:7d}{tag}")
print("tile | acc ready (abs) | drained - ready | store issued - ready | tile period")
prev = None
for k in range(min(ntl, 8)):
    a, d, s = t[1900 + 4 * k: 1900 + 4 * k + 3]
    print(f"{k:3d} | {a - t0:9d} | {d - a:6d} | {s - a:6d} | {(a - prev) if prev else 0}")
    prev = a
