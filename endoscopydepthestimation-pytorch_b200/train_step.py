"""The body of the reference train loop (`/root/reference/train.py:272-328`) on top of the mirrored
modules.  `LossStack.loss` is a line-for-line use of the public module API exactly as train.py calls
it (so it doubles as the drop-in check); `TrainStep` adds the optimiser tail (`clip_grad_norm_` +
SGD momentum) as the fused flat-bucket kernel and, on several GPUs, the gradient all-reduce.
"""
import math

import torch

from . import losses, models, optim


class LossStack:
    """train.py:206-211 (layer / loss construction) and :272-315 (per-step use)."""

    def __init__(self, height, width, dcl_weight=5.0, sfl_weight=20.0, epsilon=1.0e-8):
        self.depth_scaling_layer = models.DepthScalingLayer(epsilon=epsilon)          # train.py:206
        self.depth_warping_layer = models.DepthWarpingLayer(epsilon=epsilon)          # :207
        self.flow_from_depth_layer = models.FlowfromDepthLayer()                      # :208
        self.sparse_flow_loss_function = losses.SparseMaskedL1Loss()                  # :210
        self.depth_consistency_loss_function = losses.NormalizedDistanceLoss(height=height, width=width)  # :211
        self.depth_consistency_weight = float(dcl_weight)
        self.sparse_flow_weight = float(sfl_weight)

    def loss(self, net, batch, pair=False):
        b = batch
        boundaries = b["boundaries"]
        colors_1 = boundaries * b["colors_1"]                                         # :272-273
        colors_2 = boundaries * b["colors_2"]
        if pair:
            predicted_depth_maps_1, predicted_depth_maps_2 = net.forward_pair(colors_1, colors_2)
        else:
            predicted_depth_maps_1 = net(colors_1)                                    # :276-277
            predicted_depth_maps_2 = net(colors_2)
        scaled_depth_maps_1, std_1 = self.depth_scaling_layer(
            [predicted_depth_maps_1, b["sparse_depths_1"], b["sparse_depth_masks_1"]])   # :279-282
        scaled_depth_maps_2, std_2 = self.depth_scaling_layer(
            [predicted_depth_maps_2, b["sparse_depths_2"], b["sparse_depth_masks_2"]])
        flows_from_depth_1 = self.flow_from_depth_layer(
            [scaled_depth_maps_1, boundaries, b["translations_1_wrt_2"], b["rotations_1_wrt_2"], b["intrinsics"]])
        flows_from_depth_2 = self.flow_from_depth_layer(
            [scaled_depth_maps_2, boundaries, b["translations_2_wrt_1"], b["rotations_2_wrt_1"], b["intrinsics"]])
        sparse_flow_masks_1 = b["sparse_flow_masks_1"] * boundaries                   # :293-298
        sparse_flow_masks_2 = b["sparse_flow_masks_2"] * boundaries
        sparse_flows_1 = b["sparse_flows_1"] * boundaries
        sparse_flows_2 = b["sparse_flows_2"] * boundaries
        flows_from_depth_1 = flows_from_depth_1 * boundaries
        flows_from_depth_2 = flows_from_depth_2 * boundaries
        sparse_flow_loss = self.sparse_flow_weight * 0.5 * (
            self.sparse_flow_loss_function([sparse_flows_1, flows_from_depth_1, sparse_flow_masks_1]) +
            self.sparse_flow_loss_function([sparse_flows_2, flows_from_depth_2, sparse_flow_masks_2]))   # :300-302
        warped_depth_maps_2_to_1, intersect_masks_1 = self.depth_warping_layer(
            [scaled_depth_maps_1, scaled_depth_maps_2, boundaries, b["translations_1_wrt_2"],
             b["rotations_1_wrt_2"], b["intrinsics"]])                                # :305-310
        warped_depth_maps_1_to_2, intersect_masks_2 = self.depth_warping_layer(
            [scaled_depth_maps_2, scaled_depth_maps_1, boundaries, b["translations_2_wrt_1"],
             b["rotations_2_wrt_1"], b["intrinsics"]])
        depth_consistency_loss = self.depth_consistency_weight * 0.5 * (
            self.depth_consistency_loss_function(
                [scaled_depth_maps_1, warped_depth_maps_2_to_1, intersect_masks_1, b["intrinsics"]]) +
            self.depth_consistency_loss_function(
                [scaled_depth_maps_2, warped_depth_maps_1_to_2, intersect_masks_2, b["intrinsics"]]))   # :311-314
        loss = depth_consistency_loss + sparse_flow_loss                              # :315
        extras = dict(depth_1=predicted_depth_maps_1, depth_2=predicted_depth_maps_2, scaled_1=scaled_depth_maps_1,
                      scaled_2=scaled_depth_maps_2, warped_2to1=warped_depth_maps_2_to_1,
                      warped_1to2=warped_depth_maps_1_to_2, inter_1=intersect_masks_1, inter_2=intersect_masks_2,
                      flow_1=flows_from_depth_1, flow_2=flows_from_depth_2, std_1=std_1, std_2=std_2)
        return loss, depth_consistency_loss, sparse_flow_loss, extras

    def forward_backward(self, net, batch, pair=False):
        for p in net.parameters():
            p.grad = None                                                             # optimizer.zero_grad(), :324
        loss, dcl, sfl, extras = self.loss(net, batch, pair=pair)
        loss.backward()                                                               # :325
        return loss.detach(), dcl.detach(), sfl.detach(), extras


class TrainStep:
    """One optimisation step (train.py:272-328) with the fused optimiser tail.

    `step(batch)` returns the loss as a DEVICE tensor; the NaN/Inf guard of train.py:317-322 runs on the
    device (the update is skipped when the loss is not finite) so the step needs no host sync."""

    def __init__(self, net, height, width, lr=1.0e-3, momentum=0.9, max_norm=10.0, dcl_weight=5.0, sfl_weight=20.0,
                 epsilon=1.0e-8, pair=True, process_group=None):
        self.net = net
        self.stack = LossStack(height, width, dcl_weight, sfl_weight, epsilon)
        self.opt = optim.FusedClipSGD(net, lr=lr, momentum=momentum, max_norm=max_norm)
        self.pair = pair
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1

    def step(self, batch):
        loss, dcl, sfl, _ = self.stack.forward_backward(self.net, batch, pair=self.pair)
        finite = torch.isfinite(loss).to(torch.float32).reshape(1)
        if self.world > 1:
            from . import ddp
            ddp.allreduce_gradients(self.net, finite, self.pg)
        self.opt.step(finite_flag=finite)
        return loss, dcl, sfl


def is_bad(loss_value: float) -> bool:
    return math.isnan(loss_value) or math.isinf(loss_value)                           # train.py:317


class GraphedTrainStep:
    """The whole optimisation step (train.py:272-328) captured ONCE in a CUDA graph and replayed per step.

    The step has no host-side control flow: the NaN guard of train.py:317-322 is the device-side is-finite flag of the
    fused optimiser tail, the learning rate is a device scalar (`set_lr`, so an LR schedule does not invalidate the
    graph) and the BatchNorm / optimiser state lives in fixed flat arrays.  Inputs are copied into fixed device buffers:

        ts = GraphedTrainStep(net, h, w, example_batch)            # warm-up + capture
        loss, dcl, sfl = ts.step(device_batch)                      # D2D into the static inputs + one graph launch
        # host-resident batches (pinned memory) with the copy of step i+1 overlapping the compute of step i:
        ts.prefetch(host_batch0)
        for batch in loader:                                        # batch = the NEXT step's pinned host tensors
            ts.swap_in(); ts.prefetch(batch); loss, dcl, sfl = ts.replay(); value = loss.item()

    On several GPUs the gradient all-reduce runs between two graphs (forward + backward | optimiser tail)."""

    def __init__(self, net, height, width, example_batch, lr=1.0e-3, momentum=0.9, max_norm=10.0, dcl_weight=5.0,
                 sfl_weight=20.0, epsilon=1.0e-8, pair=True, process_group=None, warmup=3, split_graphs=None):
        from .synthetic import BATCH_KEYS_H2D
        self.keys = [k for k in BATCH_KEYS_H2D if k in example_batch]
        dev = next(net.parameters()).device
        self.dev = dev
        self.net = net
        self.inner = TrainStep(net, height, width, lr=lr, momentum=momentum, max_norm=max_norm, dcl_weight=dcl_weight,
                               sfl_weight=sfl_weight, epsilon=epsilon, pair=pair, process_group=process_group)
        self.world, self.pg = self.inner.world, process_group
        # two graphs (forward + backward | optimiser tail) around the gradient all-reduce on several GPUs; `split_graphs=True`
        # forces that shape on one GPU (tests)
        self.split = (self.world > 1) if split_graphs is None else bool(split_graphs)
        self.static = {k: example_batch[k].to(dev, copy=True) for k in self.keys}
        self.staging = {k: torch.empty_like(v) for k, v in self.static.items()}
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.copy_done = torch.cuda.Event()
        self.staging_free = torch.cuda.Event()
        self.staging_free.record(torch.cuda.current_stream(dev))
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.static.values())
        from . import _lib
        # warm-up on a side stream (allocator, lazily created library state: per-device kernel attributes, side-stream pool)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 2)):
                self._eager_fwd_bwd()
                self._eager_tail()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        n0 = _lib.launch_count()
        self.g_main = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_main):
            self._eager_fwd_bwd()
            if not self.split:
                self._eager_tail()
        self.g_tail = None
        if self.split:
            self.g_tail = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_tail):
                self._eager_tail()
        self.launches_per_step = int(_lib.launch_count() - n0)     # library kernels captured = launched by every replay

    def _eager_fwd_bwd(self):
        self.loss, self.dcl, self.sfl, _ = self.inner.stack.forward_backward(self.net, self.static, pair=self.inner.pair)
        self.finite = torch.isfinite(self.loss).to(torch.float32).reshape(1)

    def _eager_tail(self):
        self.inner.opt.step_graph(self.lr_dev, finite_flag=self.finite)

    def set_lr(self, lr):
        self.lr_dev.fill_(float(lr))

    @property
    def grad_norm(self):
        return self.inner.opt.grad_norm

    def replay(self):
        self.g_main.replay()
        if self.split:
            if self.world > 1:
                from . import ddp
                ddp.allreduce_gradients(self.net, self.finite, self.pg)
            self.g_tail.replay()
        return self.loss, self.dcl, self.sfl

    def step(self, batch):
        """`batch`: device tensors (copied into the graph's static inputs unless they already are them)."""
        for k in self.keys:
            if batch[k].data_ptr() != self.static[k].data_ptr():
                self.static[k].copy_(batch[k], non_blocking=True)
        return self.replay()

    def prefetch(self, host_batch):
        """Asynchronous H2D copy of a (pinned) host batch into the staging buffers on the copy stream."""
        self.copy_stream.wait_event(self.staging_free)
        with torch.cuda.stream(self.copy_stream):
            for k in self.keys:
                self.staging[k].copy_(host_batch[k], non_blocking=True)
            self.copy_done.record(self.copy_stream)

    def swap_in(self):
        """Make the prefetched batch the graph's input (device-to-device, ~45 MB at bs8 256x320: tens of microseconds)."""
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.copy_done)
        torch._foreach_copy_([self.static[k] for k in self.keys], [self.staging[k] for k in self.keys])
        self.staging_free.record(cur)
