"""One warm train step + N profiled steps of the bench workload (for ncu):  python tools/profile_step.py [steps] [math]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import endo_b200  # noqa: E402
from endo_b200 import train_step  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
math_mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
b, h, w = 8, 256, 320
dev = torch.device("cuda", 0)
model = endo_b200.models.FCDenseNet57(1, math=math_mode)
endo_b200.engine.kaiming_init_(model, seed=10085)
model.to(dev).train()
batch = endo_b200.synthetic.make_batch(b, h, w, seed=10085)
batch = {k: batch[k].to(dev) for k in endo_b200.synthetic.BATCH_KEYS_H2D}
ts = train_step.TrainStep(model, h, w, lr=1e-4, pair=True)
ts.step(batch)                       # warm-up (allocations, cudaFuncSetAttribute)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    loss, _, _ = ts.step(batch)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(loss))
