"""One small optimisation step (train.py:272-328) for compute-sanitizer: python tools/sanitize_step.py <math> [H W]."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import endo_b200  # noqa: E402
from endo_b200 import train_step  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (64, 96)
torch.manual_seed(1)
model = endo_b200.models.FCDenseNet57(n_classes=1, math=mode)
endo_b200.engine.kaiming_init_(model, seed=1)
with torch.no_grad():
    model.finalConv.weight.mul_(0.05)
    model.finalConv.bias.fill_(1.0)
model.cuda().train()
batch = {k: v.cuda() for k, v in endo_b200.synthetic.make_batch(2, h, w, seed=3, sparse_prob=0.02).items()}
step = train_step.TrainStep(model, h, w, lr=1e-3, pair=True)
for _ in range(2):
    loss, dcl, sfl = step.step(batch)
torch.cuda.synchronize()
print("sanitize_step", mode, h, w, "loss", float(loss), "finite grads", bool(torch.isfinite(model.flat_grads).all()))
