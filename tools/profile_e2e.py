"""Where the end-to-end arm (reference-style loop: two net() calls, torch SGD, clip_grad_norm_, loss.item()) spends its time,
next to the fused device-resident step:  python tools/profile_e2e.py"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import endo_b200
from endo_b200 import _lib, train_step

b, h, w, mode = 8, 256, 320, "tf32x3"
dev = torch.device("cuda", 0)
host = endo_b200.synthetic.make_batch(b, h, w, seed=10085)
keys = endo_b200.synthetic.BATCH_KEYS_H2D
host = {k: host[k].pin_memory() for k in keys}
model = endo_b200.models.FCDenseNet57(1, math=mode)
endo_b200.engine.kaiming_init_(model, seed=10085)
model.to(dev).train()
stack = train_step.LossStack(h, w, dcl_weight=5.0, sfl_weight=20.0)
opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.9)

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def step(marks=None):
    m = [ev()]
    cb = {k: host[k].to(dev, non_blocking=True) for k in keys}; m.append(ev())
    lv, _, _, _ = stack.loss(model, cb); m.append(ev())
    val = lv.item(); m.append(ev())
    opt.zero_grad(); lv.backward(); m.append(ev())
    torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0); opt.step(); m.append(ev())
    if marks is not None: marks.append(m)
    return val

for _ in range(3): step()
torch.cuda.synchronize()
marks = []
t0 = time.perf_counter()
for _ in range(5): step(marks)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 5
names = ["h2d", "forward+loss", "item", "backward", "clip+sgd"]
acc = [0.0] * 5
for m in marks:
    for i in range(5): acc[i] += m[i].elapsed_time(m[i + 1])
print("e2e wall ms/step", round(wall * 1e3, 2), {n: round(a / 5, 2) for n, a in zip(names, acc)})
with _lib.profile() as prof:
    for _ in range(3): step()
print("e2e kernel classes", {k: round(v / 3, 2) for k, v in prof.ms.items() if v > 0})
