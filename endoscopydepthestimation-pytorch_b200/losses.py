"""Drop-in replacements for the hot-path classes of /root/reference/losses.py.

Same class names, constructor arguments and list-packed `forward(x)` signature as the
reference, so `train.py:210-211, 300-314` runs unchanged on top of them; the arithmetic is one
fused CUDA kernel per forward and per backward (csrc/losses.cu) instead of 15-40 eager ops.
"""
import torch
from torch import nn

from . import functional as F_


class SparseMaskedL1Loss(nn.Module):
    """/root/reference/losses.py:57-66.  x = [flows, flows_from_depth, sparse_masks] -> scalar."""

    def __init__(self, epsilon=1.0):
        super().__init__()
        self.epsilon = float(epsilon)

    def forward(self, x):
        flows, flows_from_depth, sparse_masks = x
        return F_.SparseL1Fn.apply(flows, flows_from_depth, sparse_masks, self.epsilon)


class NormalizedDistanceLoss(nn.Module):
    """/root/reference/losses.py:112-146.  x = [depth_maps, warped_depth_maps, intersect_masks, intrinsics]."""

    def __init__(self, height, width, eps=1.0e-5):
        super().__init__()
        self.height, self.width, self.eps = int(height), int(width), float(eps)

    def forward(self, x):
        depth_maps, warped_depth_maps, intersect_masks, intrinsics = x
        if depth_maps.shape[2] != self.height or depth_maps.shape[3] != self.width:
            raise RuntimeError(f"NormalizedDistanceLoss was built for {self.height}x{self.width} inputs, got "
                               f"{tuple(depth_maps.shape[2:])}")     # the reference's meshgrid would fail to broadcast
        return F_.NormDistFn.apply(depth_maps, warped_depth_maps, intersect_masks, intrinsics, self.eps)


class ScaleInvariantLoss(nn.Module):
    """/root/reference/losses.py:17-32.  x = [predicted_depths, goal_depths, boundaries] -> scalar."""

    def __init__(self, epsilon=1.0e-8):
        super().__init__()
        self.epsilon = float(epsilon)

    def forward(self, x):
        predicted_depths, goal_depths, boundaries = x
        return F_.ScaleInvFn.apply(predicted_depths, goal_depths, boundaries, self.epsilon)
