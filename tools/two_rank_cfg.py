"""torchrun --nproc-per-node 2 tools/two_rank_cfg.py <bs> <H> <W> <math>: GraphedTrainStep on 2 ranks, phase by phase (debug aid)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist
import endo_b200
from endo_b200 import ddp, train_step
bs, h, w, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
rank, world, local = ddp.init_from_env("nccl")
dev = torch.device("cuda", local)
def say(*a):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    if rank == 0: print(*a, flush=True)
m = endo_b200.models.FCDenseNet57(n_classes=1, math=mode)
endo_b200.engine.kaiming_init_(m, seed=1)
with torch.no_grad():
    m.finalConv.weight.mul_(0.05); m.finalConv.bias.fill_(1.0)
m.to(dev).train()
hb = endo_b200.synthetic.make_batch(bs, h, w, seed=10085 + rank)
rb = {k: hb[k].to(dev) for k in endo_b200.synthetic.BATCH_KEYS_H2D}
say("built", bs, h, w, mode)
eager = train_step.TrainStep(m, h, w, lr=1e-4, pair=True, process_group=dist.group.WORLD)
l, _, _ = eager.step(rb); say("eager step ok", float(l))
g = train_step.GraphedTrainStep(m, h, w, rb, lr=1e-4, pair=True, process_group=dist.group.WORLD, warmup=2)
say("captured; launches per step", g.launches_per_step)
for i in range(3):
    l, _, _ = g.replay(); say("replay", i, float(l))
dist.destroy_process_group()
