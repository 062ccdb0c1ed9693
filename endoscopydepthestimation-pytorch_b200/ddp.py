"""Data parallelism: one process per GPU, replicated weights, ONE all-reduce of the flat gradient
bucket per step (replaces `torch.nn.DataParallel`, /root/reference/train.py:197).

The bucket is `net.flat_grads` (1,374,865 floats = 5.5 MB for FCDenseNet57); a rank-consistent
`isfinite(loss)` flag rides along so that the NaN guard of train.py:317-322 takes the same branch on
every rank.  BatchNorm statistics stay per rank, as under DataParallel."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun); returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def allreduce_bucket_(store: torch.Tensor, n: int, finite: torch.Tensor = None, group=None):
    """In-place mean over the group of the gradient bucket `store[:n]`; `finite` (1 float, 1.0 = this rank's
    loss is finite) becomes 1.0 only if every rank was finite.

    The flag rides in the spare slot `store[n]` of the SAME collective (as flag - 1: the sum is 0 when all
    ranks are finite, negative otherwise), so the step has exactly one launch-latency-bound collective."""
    world = dist.get_world_size(group)
    if world == 1:
        return
    if finite is not None:
        store[n:n + 1].copy_(finite - 1.0)
    dist.all_reduce(store[:n + 1], op=dist.ReduceOp.SUM, group=group)
    if finite is not None:
        finite.copy_((store[n:n + 1] > -0.5).to(finite.dtype))
    store[:n].mul_(1.0 / world)


def allreduce_gradients(net, finite: torch.Tensor = None, group=None):
    allreduce_bucket_(net._flat_grad_store, net.flat_grads.numel(), finite, group)


def broadcast_parameters(net, src=0, group=None):
    """Make every rank start from rank `src`'s weights and BN buffers (DataParallel replicates each forward)."""
    if dist.get_world_size(group) == 1:
        return
    if net.flat_params is None:
        net.materialize()                      # the flat arrays are otherwise built by the first forward
    dist.broadcast(net.flat_params, src=src, group=group)
    dist.broadcast(net._flat_buf, src=src, group=group)
    for t in (net._nbt or ()):                 # num_batches_tracked (int64 counters kept by the host wrapper)
        dist.broadcast(t, src=src, group=group)
