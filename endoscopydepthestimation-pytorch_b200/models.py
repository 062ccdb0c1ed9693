"""Drop-in replacements for the hot-path classes of /root/reference/models.py.

`FCDenseNet57` & co. keep the reference's module tree (so `state_dict()` keys, `utils.init_net`
and `torch.optim.SGD(model.parameters())` work as in train.py:191-203) but run through the
B200 engine in `engine.py`; the geometric layers are thin nn.Modules over csrc/geometry.cu.
"""
import torch
from torch import nn

from . import functional as F_


def __getattr__(name):   # the network classes live in engine.py (imported on first use)
    if name in ("FCDenseNet", "FCDenseNet57", "FCDenseNet67", "FCDenseNet103", "DenseLayer", "DenseBlock",
                "TransitionDown", "TransitionUp", "Bottleneck"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)


class DepthScalingLayer(nn.Module):
    """/root/reference/models.py:339-363.  x = [depth, sparse_depth, weighted_sparse_mask]
    -> (scaled depth [B,1,H,W], mean(scale_std / scale))."""

    def __init__(self, epsilon=1.0e-8):
        super().__init__()
        self.epsilon = float(epsilon)

    def forward(self, x):
        absolute_depth_estimations, input_sparse_depths, input_weighted_sparse_masks = x
        return F_.DepthScaleFn.apply(absolute_depth_estimations, input_sparse_depths, input_weighted_sparse_masks,
                                     self.epsilon)


class FlowfromDepthLayer(nn.Module):
    """/root/reference/models.py:366-374.  x = [depth, mask, t, R, K] -> flow [B,2,H,W]."""

    def forward(self, x):
        depth_maps_1, img_masks, translation_vectors, rotation_matrices, intrinsic_matrices = x
        return F_.FlowFromDepthFn.apply(depth_maps_1, img_masks, translation_vectors, rotation_matrices,
                                        intrinsic_matrices)


class DepthWarpingLayer(nn.Module):
    """/root/reference/models.py:454-465.  x = [depth_1, depth_2, mask, t, R, K]
    -> (warped depth map 2 in frame 1, intersect mask)."""

    def __init__(self, epsilon=1.0e-8):
        super().__init__()
        self.epsilon = float(epsilon)

    def forward(self, x):
        depth_maps_1, depth_maps_2, img_masks, translation_vectors, rotation_matrices, intrinsic_matrices = x
        return F_.DepthWarpFn.apply(depth_maps_1, depth_maps_2, img_masks, translation_vectors, rotation_matrices,
                                    intrinsic_matrices, self.epsilon)
