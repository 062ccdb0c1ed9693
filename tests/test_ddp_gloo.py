"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo run of the single gradient all-reduce with
the piggy-backed is-finite flag (replaces DataParallel, train.py:197)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import endo_b200  # noqa: F401  (sets up the import path for spawned workers)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _FakeNet:
    """Stands in for FCDenseNet after its first backward: a flat bucket with 4 spare floats behind it."""

    def __init__(self, n, rank):
        self._flat_grad_store = torch.zeros(n + 4)
        self._flat_grad_store[:n] = torch.arange(n, dtype=torch.float32) * (rank + 1)
        self.flat_grads = self._flat_grad_store[:n]
        self.flat_params = torch.full((n,), float(rank))
        self._flat_buf = torch.full((7,), float(rank) + 10.0)
        self._nbt = [torch.tensor(rank + 5, dtype=torch.long)]


def _worker(rank, world, port, results):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from endo_b200 import ddp
    r, w, _ = ddp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    n = 1000
    net = _FakeNet(n, rank)
    ddp.broadcast_parameters(net, src=0)
    assert float(net.flat_params[0]) == 0.0 and float(net._flat_buf[0]) == 10.0 and int(net._nbt[0]) == 5
    # step 1: every rank finite -> flag stays 1, gradients are averaged
    finite = torch.ones(1)
    ddp.allreduce_gradients(net, finite)
    expect = torch.arange(n, dtype=torch.float32) * (1 + 2) / 2.0
    ok1 = torch.allclose(net.flat_grads, expect) and float(finite) == 1.0
    # step 2: rank 1 saw a NaN loss -> every rank must take the skip branch (train.py:317-322 made rank-consistent)
    net2 = _FakeNet(n, rank)
    finite = torch.tensor([0.0 if rank == 1 else 1.0])
    ddp.allreduce_gradients(net2, finite)
    ok2 = float(finite) == 0.0
    # no flag: plain average
    net3 = _FakeNet(n, rank)
    ddp.allreduce_gradients(net3, None)
    ok3 = torch.allclose(net3.flat_grads, expect)
    results[rank] = bool(ok1 and ok2 and ok3)
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert results[0] and results[1]


def test_single_process_is_a_noop():
    from endo_b200 import ddp
    if dist.is_initialized():
        dist.destroy_process_group()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        net = _FakeNet(10, 0)
        before = net.flat_grads.clone()
        ddp.allreduce_gradients(net, torch.ones(1))
        assert torch.equal(net.flat_grads, before)
    finally:
        dist.destroy_process_group()
