"""autograd.Function bindings of the geometric layers and losses to the C ABI.

Each Function's forward / backward is one call into libendo_b200.so on the current CUDA
stream; tensors are allocated by PyTorch and passed as raw device pointers.  Nothing here
computes on the host or falls back to PyTorch ops.
"""
import torch

from . import _lib as L


def _dims(t, channels=1, what="input"):
    if t.dim() != 4 or t.shape[1] != channels:
        raise RuntimeError(f"{what}: expected a [B,{channels},H,W] tensor, got {tuple(t.shape)}")
    b, c, h, w = t.shape
    return b, h, w


def _pose(t, r, k, b, what):
    L.require_cuda(t, r, k)
    L.same_shape(what, t, ("translation_vectors", t, (b, 3, 1)), ("rotation_matrices", r, (b, 3, 3)),
                 ("intrinsic_matrices", k, (b, 3, 3)))
    return L.contig(t), L.contig(r), L.contig(k)


class DepthScaleFn(torch.autograd.Function):
    """DepthScalingLayer.forward (/root/reference/models.py:346-363)."""

    @staticmethod
    @L.on_device
    def forward(ctx, depth, sparse_depth, sparse_mask, epsilon):
        L.require_cuda(depth, sparse_depth, sparse_mask)
        b, h, w = _dims(depth, 1, "DepthScalingLayer")
        L.same_shape("DepthScalingLayer", depth, ("sparse_depths", sparse_depth, depth.shape), ("sparse_masks", sparse_mask, depth.shape))
        depth, sparse_depth, sparse_mask = L.contig(depth), L.contig(sparse_depth), L.contig(sparse_mask)
        lib = L.lib()
        scaled = torch.empty_like(depth)
        norm_std = torch.empty((), dtype=torch.float32, device=depth.device)
        stats = torch.empty(b * 4, dtype=torch.float32, device=depth.device)
        nbytes = lib.endo_depth_scale_workspace_bytes(b, h, w)
        ws = L.workspace(depth.device, nbytes)
        L.check(lib.endo_depth_scale_fwd(depth.data_ptr(), sparse_depth.data_ptr(), sparse_mask.data_ptr(),
                                         scaled.data_ptr(), norm_std.data_ptr(), stats.data_ptr(), b, h, w,
                                         float(epsilon), ws.data_ptr(), ws.numel(), L.stream_ptr(depth.device)),
                "depth_scale_fwd")
        ctx.save_for_backward(depth, sparse_depth, stats)
        ctx.eps = float(epsilon)
        ctx.mark_non_differentiable(norm_std)
        return scaled, norm_std

    @staticmethod
    @L.on_device
    def backward(ctx, g_scaled, _g_std):
        depth, sparse_depth, stats = ctx.saved_tensors
        b, h, w = _dims(depth)
        lib = L.lib()
        g_scaled = L.contig(g_scaled)
        g_depth = torch.empty_like(depth)
        nbytes = lib.endo_depth_scale_workspace_bytes(b, h, w)
        ws = L.workspace(depth.device, nbytes)
        L.check(lib.endo_depth_scale_bwd(g_scaled.data_ptr(), depth.data_ptr(), sparse_depth.data_ptr(),
                                         stats.data_ptr(), g_depth.data_ptr(), b, h, w, ctx.eps, ws.data_ptr(),
                                         ws.numel(), L.stream_ptr(depth.device)), "depth_scale_bwd")
        return g_depth, None, None, None


class FlowFromDepthFn(torch.autograd.Function):
    """FlowfromDepthLayer.forward (/root/reference/models.py:370-374, :377-451)."""

    @staticmethod
    @L.on_device
    def forward(ctx, depth, mask, t, r, k):
        L.require_cuda(depth, mask)
        b, h, w = _dims(depth, 1, "FlowfromDepthLayer")
        L.same_shape("FlowfromDepthLayer", depth, ("img_masks", mask, depth.shape))
        depth, mask = L.contig(depth), L.contig(mask)
        t, r, k = _pose(t, r, k, b, "FlowfromDepthLayer")
        flow = torch.empty((b, 2, h, w), dtype=torch.float32, device=depth.device)
        L.check(L.lib().endo_flow_from_depth_fwd(depth.data_ptr(), mask.data_ptr(), t.data_ptr(), r.data_ptr(),
                                                 k.data_ptr(), flow.data_ptr(), b, h, w, L.stream_ptr(depth.device)),
                "flow_from_depth_fwd")
        ctx.save_for_backward(depth, mask, t, r, k)
        return flow

    @staticmethod
    @L.on_device
    def backward(ctx, g_flow):
        depth, mask, t, r, k = ctx.saved_tensors
        b, h, w = _dims(depth)
        g_flow = L.contig(g_flow)
        g_depth = torch.empty_like(depth)
        L.check(L.lib().endo_flow_from_depth_bwd(g_flow.data_ptr(), depth.data_ptr(), mask.data_ptr(), t.data_ptr(),
                                                 r.data_ptr(), k.data_ptr(), g_depth.data_ptr(), b, h, w,
                                                 L.stream_ptr(depth.device)), "flow_from_depth_bwd")
        return g_depth, None, None, None, None


class DepthWarpFn(torch.autograd.Function):
    """DepthWarpingLayer.forward (/root/reference/models.py:460-465, :469-554)."""

    @staticmethod
    @L.on_device
    def forward(ctx, d1, d2, mask, t, r, k, epsilon):
        L.require_cuda(d1, d2, mask)
        b, h, w = _dims(d1, 1, "DepthWarpingLayer")
        L.same_shape("DepthWarpingLayer", d1, ("depth_maps_2", d2, d1.shape), ("img_masks", mask, d1.shape))
        d1, d2, mask = L.contig(d1), L.contig(d2), L.contig(mask)
        t, r, k = _pose(t, r, k, b, "DepthWarpingLayer")
        warped = torch.empty_like(d1)
        intersect = torch.empty_like(d1)
        L.check(L.lib().endo_depth_warp_fwd(d1.data_ptr(), d2.data_ptr(), mask.data_ptr(), t.data_ptr(), r.data_ptr(),
                                            k.data_ptr(), warped.data_ptr(), intersect.data_ptr(), b, h, w,
                                            float(epsilon), L.stream_ptr(d1.device)), "depth_warp_fwd")
        ctx.save_for_backward(d1, d2, mask, t, r, k)
        ctx.eps = float(epsilon)
        ctx.mark_non_differentiable(intersect)
        return warped, intersect

    @staticmethod
    @L.on_device
    def backward(ctx, g_warped, _g_inter):
        d1, d2, mask, t, r, k = ctx.saved_tensors
        b, h, w = _dims(d1)
        g_warped = L.contig(g_warped)
        g_d1 = torch.empty_like(d1)
        g_d2 = torch.empty_like(d2)
        L.check(L.lib().endo_depth_warp_bwd(g_warped.data_ptr(), d1.data_ptr(), d2.data_ptr(), mask.data_ptr(),
                                            t.data_ptr(), r.data_ptr(), k.data_ptr(), g_d1.data_ptr(), g_d2.data_ptr(),
                                            b, h, w, ctx.eps, L.stream_ptr(d1.device)), "depth_warp_bwd")
        return g_d1, g_d2, None, None, None, None, None


def _loss_ws(t, b, h, w):
    nbytes = L.lib().endo_loss_workspace_bytes(b, h, w)
    return L.workspace(t.device, nbytes)


class SparseL1Fn(torch.autograd.Function):
    """SparseMaskedL1Loss.forward (/root/reference/losses.py:62-66)."""

    @staticmethod
    @L.on_device
    def forward(ctx, flows, flows_from_depth, masks, epsilon):
        L.require_cuda(flows, flows_from_depth, masks)
        b, h, w = _dims(masks, 1, "SparseMaskedL1Loss")
        L.same_shape("SparseMaskedL1Loss", masks, ("flows", flows, (b, 2, h, w)), ("flows_from_depth", flows_from_depth, (b, 2, h, w)))
        flows, flows_from_depth, masks = L.contig(flows), L.contig(flows_from_depth), L.contig(masks)
        loss = torch.empty((), dtype=torch.float32, device=flows.device)
        stats = torch.empty(b * 2, dtype=torch.float32, device=flows.device)
        ws = _loss_ws(flows, b, h, w)
        L.check(L.lib().endo_sparse_l1_fwd(flows.data_ptr(), flows_from_depth.data_ptr(), masks.data_ptr(),
                                           loss.data_ptr(), stats.data_ptr(), b, h, w, float(epsilon), ws.data_ptr(),
                                           ws.numel(), L.stream_ptr(flows.device)), "sparse_l1_fwd")
        ctx.save_for_backward(flows, flows_from_depth, masks, stats)
        ctx.eps = float(epsilon)
        return loss

    @staticmethod
    @L.on_device
    def backward(ctx, g_loss):
        flows, ffd, masks, stats = ctx.saved_tensors
        b, h, w = masks.shape[0], masks.shape[2], masks.shape[3]
        g_loss = L.contig(g_loss)
        g_ffd = torch.empty_like(ffd)
        g_f = torch.empty_like(flows) if ctx.needs_input_grad[0] else None
        L.check(L.lib().endo_sparse_l1_bwd(g_loss.data_ptr(), flows.data_ptr(), ffd.data_ptr(), masks.data_ptr(),
                                           stats.data_ptr(), g_ffd.data_ptr(), L.ptr(g_f), b, h, w, ctx.eps,
                                           L.stream_ptr(flows.device)), "sparse_l1_bwd")
        return g_f, g_ffd, None, None


class NormDistFn(torch.autograd.Function):
    """NormalizedDistanceLoss.forward (/root/reference/losses.py:122-146)."""

    @staticmethod
    @L.on_device
    def forward(ctx, depth, warped, intersect, intrinsics, eps):
        L.require_cuda(depth, warped, intersect, intrinsics)
        b, h, w = _dims(depth, 1, "NormalizedDistanceLoss")
        L.same_shape("NormalizedDistanceLoss", depth, ("warped_depth_maps", warped, depth.shape),
                     ("intersect_masks", intersect, depth.shape), ("intrinsics", intrinsics, (b, 3, 3)))
        depth, warped, intersect, intrinsics = (L.contig(depth), L.contig(warped), L.contig(intersect),
                                                L.contig(intrinsics))
        loss = torch.empty((), dtype=torch.float32, device=depth.device)
        stats = torch.empty(b * 4, dtype=torch.float32, device=depth.device)
        ws = _loss_ws(depth, b, h, w)
        L.check(L.lib().endo_norm_dist_fwd(depth.data_ptr(), warped.data_ptr(), intersect.data_ptr(),
                                           intrinsics.data_ptr(), loss.data_ptr(), stats.data_ptr(), b, h, w,
                                           float(eps), ws.data_ptr(), ws.numel(), L.stream_ptr(depth.device)),
                "norm_dist_fwd")
        ctx.save_for_backward(depth, warped, intersect, intrinsics, stats)
        ctx.eps = float(eps)
        return loss

    @staticmethod
    @L.on_device
    def backward(ctx, g_loss):
        depth, warped, intersect, intrinsics, stats = ctx.saved_tensors
        b, h, w = _dims(depth)
        g_loss = L.contig(g_loss)
        g_d = torch.empty_like(depth)
        g_w = torch.empty_like(warped)
        L.check(L.lib().endo_norm_dist_bwd(g_loss.data_ptr(), depth.data_ptr(), warped.data_ptr(),
                                           intersect.data_ptr(), intrinsics.data_ptr(), stats.data_ptr(),
                                           g_d.data_ptr(), g_w.data_ptr(), b, h, w, ctx.eps,
                                           L.stream_ptr(depth.device)), "norm_dist_bwd")
        return g_d, g_w, None, None, None


class ScaleInvFn(torch.autograd.Function):
    """ScaleInvariantLoss.forward (/root/reference/losses.py:22-32)."""

    @staticmethod
    @L.on_device
    def forward(ctx, pred, goal, boundaries, epsilon):
        L.require_cuda(pred, goal, boundaries)
        b, h, w = _dims(pred, 1, "ScaleInvariantLoss")
        L.same_shape("ScaleInvariantLoss", pred, ("goal_depths", goal, pred.shape), ("boundaries", boundaries, pred.shape))
        pred, goal, boundaries = L.contig(pred), L.contig(goal), L.contig(boundaries)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        stats = torch.empty(b * 4, dtype=torch.float32, device=pred.device)
        ws = _loss_ws(pred, b, h, w)
        L.check(L.lib().endo_scale_inv_fwd(pred.data_ptr(), goal.data_ptr(), boundaries.data_ptr(), loss.data_ptr(),
                                           stats.data_ptr(), b, h, w, float(epsilon), ws.data_ptr(), ws.numel(),
                                           L.stream_ptr(pred.device)), "scale_inv_fwd")
        ctx.save_for_backward(pred, goal, boundaries, stats)
        ctx.eps = float(epsilon)
        return loss

    @staticmethod
    @L.on_device
    def backward(ctx, g_loss):
        pred, goal, boundaries, stats = ctx.saved_tensors
        b, h, w = _dims(pred)
        g_loss = L.contig(g_loss)
        g_p = torch.empty_like(pred)
        g_g = torch.empty_like(goal) if ctx.needs_input_grad[1] else None
        L.check(L.lib().endo_scale_inv_bwd(g_loss.data_ptr(), pred.data_ptr(), goal.data_ptr(), boundaries.data_ptr(),
                                           stats.data_ptr(), g_p.data_ptr(), L.ptr(g_g), b, h, w, ctx.eps,
                                           L.stream_ptr(pred.device)), "scale_inv_bwd")
        return g_p, g_g, None, None
