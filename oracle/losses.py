"""Oracle restatement of the training losses (TEST INFRASTRUCTURE ONLY).

Restates /root/reference/losses.py:17-32 (ScaleInvariantLoss), :57-66 (SparseMaskedL1Loss)
and :112-146 (NormalizedDistanceLoss).  torch CPU, any float dtype, autograd gradients.
"""
import torch


def scale_invariant_loss(predicted, goal, boundaries, epsilon=1.0e-8):
    """losses.py:22-32."""
    r = torch.log(boundaries * predicted + epsilon) - torch.log(boundaries * goal + epsilon)
    wsum = boundaries.sum(dim=(1, 2, 3))
    loss_1 = (r * r).sum(dim=(1, 2, 3)) / wsum
    sum_2 = r.sum(dim=(1, 2, 3))
    loss_2 = (sum_2 * sum_2) / (wsum * wsum)
    return torch.mean(loss_1 + loss_2)


def sparse_masked_l1_loss(flows, flows_from_depth, sparse_masks, epsilon=1.0):
    """losses.py:62-66: masks [B,1,H,W] broadcast over the 2 flow channels in the numerator only."""
    loss = (sparse_masks * torch.abs(flows - flows_from_depth)).sum(dim=(1, 2, 3)) / \
           (epsilon + sparse_masks.sum(dim=(1, 2, 3)))
    return torch.mean(loss)


def normalized_distance_loss(depth, warped, intersect, intrinsics, eps=1.0e-5):
    """losses.py:122-146."""
    b, _, h, w = depth.shape
    y_grid = torch.arange(h, dtype=depth.dtype, device=depth.device).reshape(1, 1, h, 1).expand(1, 1, h, w)   # :116-120
    x_grid = torch.arange(w, dtype=depth.dtype, device=depth.device).reshape(1, 1, 1, w).expand(1, 1, h, w)
    fx = intrinsics[:, 0, 0].reshape(-1, 1, 1, 1)
    fy = intrinsics[:, 1, 1].reshape(-1, 1, 1, 1)
    cx = intrinsics[:, 0, 2].reshape(-1, 1, 1, 1)
    cy = intrinsics[:, 1, 2].reshape(-1, 1, 1, 1)
    with torch.no_grad():                                                                # :129-132
        mean_value = (intersect * depth).sum(dim=(1, 2, 3)) / (eps + intersect.sum(dim=(1, 2, 3)))
    loc = torch.cat([(x_grid - cx) / fx * depth, (y_grid - cy) / fy * depth, depth], dim=1)
    loc_w = torch.cat([(x_grid - cx) / fx * warped, (y_grid - cy) / fy * warped, warped], dim=1)
    loss = 2.0 * (intersect * torch.abs(loc - loc_w)).sum(dim=(1, 2, 3)) / \
        (1.0e-5 * mean_value + (intersect * (depth + torch.abs(warped))).sum(dim=(1, 2, 3)))
    return torch.mean(loss)
