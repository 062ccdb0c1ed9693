"""Oracle restatement of one optimisation step of the reference train loop
(TEST INFRASTRUCTURE ONLY): /root/reference/train.py:272-328.

`batch` is a dict with the tensors the reference's DataLoader yields (train.py:244-248):
colors_1, colors_2 [B,3,H,W]; sparse_depths_{1,2}, sparse_depth_masks_{1,2},
sparse_flow_masks_{1,2}, boundaries [B,1,H,W]; sparse_flows_{1,2} [B,2,H,W];
rotations_1_wrt_2, rotations_2_wrt_1 [B,3,3]; translations_1_wrt_2, translations_2_wrt_1 [B,3,1];
intrinsics [B,3,3].
"""
from collections import OrderedDict
from typing import Dict

import torch

from . import geometry, losses, net


def loss_stack(depth_1, depth_2, batch, dcl_weight=5.0, sfl_weight=20.0, epsilon=1.0e-8):
    """train.py:279-315 given the two predicted depth maps. Returns (loss, dcl, sfl, extras)."""
    b = batch
    bound = b["boundaries"]
    scaled_1, std_1 = geometry.depth_scaling(depth_1, b["sparse_depths_1"], b["sparse_depth_masks_1"], epsilon)
    scaled_2, std_2 = geometry.depth_scaling(depth_2, b["sparse_depths_2"], b["sparse_depth_masks_2"], epsilon)
    flow_1 = geometry.flow_from_depth(scaled_1, bound, b["translations_1_wrt_2"], b["rotations_1_wrt_2"],
                                      b["intrinsics"])
    flow_2 = geometry.flow_from_depth(scaled_2, bound, b["translations_2_wrt_1"], b["rotations_2_wrt_1"],
                                      b["intrinsics"])
    sfm_1 = b["sparse_flow_masks_1"] * bound                                   # train.py:293-298
    sfm_2 = b["sparse_flow_masks_2"] * bound
    sf_1 = b["sparse_flows_1"] * bound
    sf_2 = b["sparse_flows_2"] * bound
    flow_1 = flow_1 * bound
    flow_2 = flow_2 * bound
    sfl = sfl_weight * 0.5 * (losses.sparse_masked_l1_loss(sf_1, flow_1, sfm_1) +
                              losses.sparse_masked_l1_loss(sf_2, flow_2, sfm_2))  # :300-302
    warped_2to1, inter_1 = geometry.depth_warping(scaled_1, scaled_2, bound, b["translations_1_wrt_2"],
                                                  b["rotations_1_wrt_2"], b["intrinsics"], epsilon)
    warped_1to2, inter_2 = geometry.depth_warping(scaled_2, scaled_1, bound, b["translations_2_wrt_1"],
                                                  b["rotations_2_wrt_1"], b["intrinsics"], epsilon)
    dcl = dcl_weight * 0.5 * (losses.normalized_distance_loss(scaled_1, warped_2to1, inter_1, b["intrinsics"]) +
                              losses.normalized_distance_loss(scaled_2, warped_1to2, inter_2, b["intrinsics"]))
    loss = dcl + sfl                                                            # :315
    extras = dict(scaled_1=scaled_1, scaled_2=scaled_2, flow_1=flow_1, flow_2=flow_2,
                  warped_2to1=warped_2to1, warped_1to2=warped_1to2, inter_1=inter_1, inter_2=inter_2,
                  std_1=std_1, std_2=std_2)
    return loss, dcl, sfl, extras


def forward_backward(state: Dict[str, torch.Tensor], batch, cfg=net.FCDENSENET57, dcl_weight=5.0,
                     sfl_weight=20.0, epsilon=1.0e-8):
    """train.py:272-325: returns (loss, dcl, sfl, grads{name}, new BN buffers, extras)."""
    params = OrderedDict()
    for k, v in state.items():
        if net.is_buffer(k):
            params[k] = v
        else:
            params[k] = v.detach().clone().requires_grad_(True)
    new_buffers: Dict[str, torch.Tensor] = {}
    colors_1 = batch["boundaries"] * batch["colors_1"]                          # :272-273
    colors_2 = batch["boundaries"] * batch["colors_2"]
    depth_1 = net.forward(params, colors_1, cfg, True, new_buffers)             # :276-277 (two BN-train forwards)
    depth_2 = net.forward(params, colors_2, cfg, True, new_buffers)
    loss, dcl, sfl, extras = loss_stack(depth_1, depth_2, batch, dcl_weight, sfl_weight, epsilon)
    loss.backward()                                                             # :325
    grads = OrderedDict((k, p.grad) for k, p in params.items() if not net.is_buffer(k))
    extras.update(depth_1=depth_1.detach(), depth_2=depth_2.detach())
    return loss.detach(), dcl.detach(), sfl.detach(), grads, new_buffers, extras


def clip_and_sgd(state, grads, momentum_buffers, lr, max_norm=10.0, momentum=0.9):
    """`clip_grad_norm_(…, 10.0)` + `SGD(momentum=0.9).step()` (train.py:202, 327-328).

    Updates `state` and `momentum_buffers` in place; returns the total gradient norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).to(next(iter(grads.values())).dtype)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)                      # torch.nn.utils.clip_grad_norm_
    with torch.no_grad():
        for k, g in grads.items():
            g = g * coef
            if k not in momentum_buffers:
                momentum_buffers[k] = g.clone()                                 # first step: buf = grad
            else:
                momentum_buffers[k].mul_(momentum).add_(g)
            state[k] = state[k] - lr * momentum_buffers[k]
    return total
