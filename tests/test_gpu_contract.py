"""The nn.Module contract of the engine where it deviates from eager PyTorch: deviations must be loud errors
(ADVICE r1), the flat arrays can be materialised before the first forward (ddp.broadcast_parameters), shape
mismatches raise instead of reading out of bounds."""
import pytest
import torch

import endo_b200

pytestmark = pytest.mark.gpu


def test_materialize_and_broadcast_before_first_forward():
    import os
    import socket
    import torch.distributed as dist
    from endo_b200 import ddp
    model = endo_b200.models.FCDenseNet57(n_classes=1).cuda()
    assert model.flat_params is None
    model.materialize()
    n = sum(p.numel() for p in model.parameters())
    assert model.flat_params.numel() == n and model.flat_params.is_cuda
    assert next(model.parameters()).data_ptr() == model.flat_params.data_ptr()
    # single-process group: the documented call must work on a real engine module before its first step
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if dist.is_initialized():
        dist.destroy_process_group()
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        fresh = endo_b200.models.FCDenseNet57(n_classes=1).cuda()
        ddp.broadcast_parameters(fresh, src=0)
    finally:
        dist.destroy_process_group()


def test_deviations_from_autograd_contract_raise():
    model = endo_b200.models.FCDenseNet57(n_classes=1).cuda().train()
    x = torch.rand(2, 3, 64, 64, device="cuda")
    y = model(x)
    y.sum().backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second backward"):
        y.sum().backward()
    y = model(x)
    with torch.no_grad():
        model.flat_params.mul_(1.0)                       # in-place update between forward and backward
    with pytest.raises(RuntimeError, match="modified in place"):
        y.sum().backward()
    model.firstconv.weight.requires_grad_(False)
    with pytest.raises(RuntimeError, match="frozen parameters"):
        model(x)
    with torch.no_grad():
        model(x)                                          # inference with frozen parameters is fine


def test_operand_shape_mismatch_raises():
    b, h, w = 2, 32, 64
    d = torch.rand(b, 1, h, w, device="cuda") + 0.5
    m = torch.ones(b, 1, h, w, device="cuda")
    t = torch.zeros(b, 3, 1, device="cuda"); r = torch.eye(3, device="cuda").repeat(b, 1, 1); k = r.clone()
    warp = endo_b200.models.DepthWarpingLayer()
    with pytest.raises(RuntimeError, match="img_masks"):
        warp([d, d, m[:, :, :16], t, r, k])
    with pytest.raises(RuntimeError, match="rotation_matrices"):
        warp([d, d, m, t, r[:1], k])
    flow = endo_b200.models.FlowfromDepthLayer()
    with pytest.raises(RuntimeError, match="translation_vectors"):
        flow([d, m, t.reshape(b, 3), r, k])
    l1 = endo_b200.losses.SparseMaskedL1Loss()
    with pytest.raises(RuntimeError, match="flows"):
        l1([d, d, m])                                     # flows must be [B,2,H,W]
    ndl = endo_b200.losses.NormalizedDistanceLoss(height=h, width=w)
    with pytest.raises(RuntimeError, match="intersect_masks"):
        ndl([d, d, m[:1], k])
