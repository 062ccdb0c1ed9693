"""Why the 25-step bench loss of round 1 differed between identical runs: CPU-ORACLE experiment, no GPU.

Runs N optimisation steps of train.py:272-328 on the fp32 CPU oracle twice -- once from the seeded state, once with ONE
weight changed by 1 ulp -- for (a) the raw Kaiming initialisation round 1 benchmarked and (b) the well-conditioned start the
round-2 bench and fixtures use (oracle.net.condition_state).  The CPU oracle is deterministic, so the divergence it shows is
a property of the loss landscape at that initialisation, not of any kernel: a 1-ulp perturbation plays the role of the
summation-order noise of fp32 atomics (run-to-run gradient differences of 3e-6 of the max on the GPU).

    python tools/chaos_probe.py [steps] [H W]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import endo_b200  # noqa: E402
from oracle import net as onet, step as ostep  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (64, 64)
torch.set_num_threads(8)
cfg = onet.FCDENSENET57
batch = endo_b200.synthetic.make_batch(2, h, w, seed=10085, sparse_prob=0.02)


def run(state):
    state = dict(state)
    mom, out = {}, []
    for _ in range(steps):
        loss, _, _, grads, new_buf, _ = ostep.forward_backward(state, batch, cfg, 5.0, 20.0)
        gn = ostep.clip_and_sgd(state, grads, mom, lr=1e-4)
        state.update(new_buf)
        out.append((float(loss), float(gn)))
    return out


for name, cond in (("raw Kaiming init (round-1 bench)", False), ("conditioned finalConv (round-2 bench)", True)):
    base = onet.init_state(cfg, seed=10085)
    if cond:
        base = onet.condition_state(base)
    pert = dict(base)
    wt = base["denseBlocksDown.0.layers.0.conv.weight"].clone()
    wt.view(-1)[0] = float(np.nextafter(np.float32(float(wt.view(-1)[0])), np.float32(np.inf)))      # +1 ulp on one weight
    pert["denseBlocksDown.0.layers.0.conv.weight"] = wt
    a, b = run(base), run(pert)
    print(name)
    for i in (0, 1, 2, steps // 2, steps - 1):
        la, lb = a[i][0], b[i][0]
        print(f"  step {i + 1:2d}: loss {la:.6f} vs {lb:.6f}  rel diff {abs(la - lb) / abs(la):.2e}   grad norm {a[i][1]:.3e}")
