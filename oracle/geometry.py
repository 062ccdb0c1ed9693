"""Oracle restatement of the differentiable geometric layers (TEST INFRASTRUCTURE ONLY).

Restates /root/reference/models.py:325-554 (`_bilinear_interpolate`, `DepthScalingLayer`,
`FlowfromDepthLayer`, `DepthWarpingLayer`).  The bilinear sampler is written out tap by tap
instead of calling `F.grid_sample`: the reference calls `grid_sample` without
`align_corners` (models.py:335) on a grid normalised as 2*x/W-1 (models.py:328-333), which on
the installed torch means sampling at pixel coordinates (u-0.5, v-0.5) with zeros padding
(SURVEY.md section 0).  Tensors are torch CPU (or, for bench.py's reference-on-GPU arm, CUDA) tensors, any float dtype, NCHW like the reference's API.
"""
import torch


def _pose_terms(translation, rotation, intrinsics):
    """models.py:391-399 / 492-499: K^-1 (solve K X = I), T = K R^T, W = T(-t), M = T K^-1."""
    b = intrinsics.shape[0]
    eye = torch.eye(3, dtype=intrinsics.dtype, device=intrinsics.device).reshape(1, 3, 3).expand(b, -1, -1)
    k_inv = torch.linalg.solve(intrinsics, eye)
    temp = torch.bmm(intrinsics, rotation.transpose(1, 2))
    w_vec = torch.bmm(temp, -translation.reshape(b, 3, 1)).reshape(b, 3)
    m_mat = torch.bmm(temp, k_inv)
    return k_inv, w_vec, m_mat


def _mesh(height, width, dtype, device=None):
    """models.py:381-386: meshgrid 'ij' => x_grid[h, w] = w, y_grid[h, w] = h."""
    y = torch.arange(height, dtype=dtype, device=device).reshape(1, 1, height, 1).expand(1, 1, height, width)
    x = torch.arange(width, dtype=dtype, device=device).reshape(1, 1, 1, width).expand(1, 1, height, width)
    return x, y


def depth_scaling(depth, sparse_depth, sparse_mask, epsilon=1.0e-8):
    """`DepthScalingLayer.forward` (models.py:346-363) -> (scaled depth, mean(std / scale))."""
    one = torch.ones((), dtype=depth.dtype)
    zero = torch.zeros((), dtype=depth.dtype)
    bm = torch.where(sparse_mask > 1.0e-8, one, zero)                         # :350
    mean_sd = (sparse_depth * bm).sum(dim=(1, 2, 3), keepdim=True) / bm.sum(dim=(1, 2, 3), keepdim=True)
    am = torch.where(sparse_depth > 0.5 * mean_sd, one, zero)                 # :353
    scale_map = sparse_depth * am / (epsilon + depth)                         # :356
    n_am = am.sum(dim=(1, 2, 3), keepdim=True)
    mean_scale = scale_map.sum(dim=(1, 2, 3), keepdim=True) / n_am            # :357
    centered = scale_map - am * mean_scale                                    # :359
    std = torch.sqrt((centered * centered).sum(dim=(1, 2, 3)) / am.sum(dim=(1, 2, 3)))   # :360
    scales = scale_map.sum(dim=(1, 2, 3)) / am.sum(dim=(1, 2, 3))             # :362
    return scales.reshape(-1, 1, 1, 1) * depth, torch.mean(std / mean_scale)  # :363 (broadcast [B]/[B,1,1,1])


def flow_from_depth(depth, mask, translation, rotation, intrinsics):
    """`FlowfromDepthLayer.forward` (models.py:370-374 -> :433-451 -> :377-429). Returns [B,2,H,W]."""
    b, _, h, w = depth.shape
    _, w_vec, m_mat = _pose_terms(translation, rotation, intrinsics)
    x, y = _mesh(h, w, depth.dtype, depth.device)
    m = m_mat.reshape(b, 3, 3, 1, 1)
    q = [m[:, r, 0] * x[0] + m[:, r, 1] * y[0] + m[:, r, 2] for r in range(3)]       # M . [x, y, 1]  (:401-402)
    q = [t.reshape(b, 1, h, w) for t in q]
    wv = w_vec.reshape(b, 3, 1, 1, 1)
    z2 = wv[:, 2] + depth * q[2]                                                # :404-407
    z2 = 1.0e30 * (1.0 - mask) + mask * z2                                      # :410-411
    u2 = (wv[:, 0] + depth * q[0]) / z2                                         # :414-420
    v2 = (wv[:, 1] + depth * q[1]) / z2                                         # :422-428
    return torch.cat([(u2 - x) / float(w), (v2 - y) / float(h)], dim=1)          # :449-451


LIBRARY_OPS = False   # see oracle/net.py: call F.grid_sample like the reference (CPU-baseline timing only)


def _bilinear_library(src, u, v):
    b, _, h, w = src.shape
    grid = torch.cat([(2.0 * (u.reshape(b, h, w, 1) / float(w)) - 1.0), (2.0 * (v.reshape(b, h, w, 1) / float(h)) - 1.0)],
                     dim=-1)
    return torch.nn.functional.grid_sample(src, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


def bilinear_zero_pad(src, u, v):
    """`_bilinear_interpolate` (models.py:325-336) restated on pixel coordinates.

    src [B,1,H,W]; u, v [B,1,H,W] in the reference's convention (grid = 2u/W-1 fed to
    grid_sample(align_corners=False, zeros)) => sample location ix = u-0.5, iy = v-0.5, four
    taps, a tap contributes only when it lies inside the image.
    """
    if LIBRARY_OPS:
        return _bilinear_library(src, u, v)
    b, _, h, w = src.shape
    # grid_sample un-normalisation: ix = ((g + 1) * W - 1) / 2 with g = 2u/W - 1
    gx = 2.0 * (u / float(w)) - 1.0
    gy = 2.0 * (v / float(h)) - 1.0
    ix = ((gx + 1.0) * w - 1.0) / 2.0
    iy = ((gy + 1.0) * h - 1.0) / 2.0
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1 = x0 + 1.0
    y1 = y0 + 1.0
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    flat = src.reshape(b, h * w)

    def tap(xi, yi, wt):
        valid = (xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)
        # clamp in floating point first: coordinates can be astronomically large (division by epsilon)
        xc = torch.nan_to_num(xi, nan=0.0).clamp(0, w - 1).long()
        yc = torch.nan_to_num(yi, nan=0.0).clamp(0, h - 1).long()
        idx = (yc * w + xc).reshape(b, h * w)
        val = torch.gather(flat, 1, idx).reshape(b, 1, h, w)
        return torch.where(valid, val * wt, torch.zeros((), dtype=src.dtype))

    return tap(x0, y0, w_nw) + tap(x1, y0, w_ne) + tap(x0, y1, w_sw) + tap(x1, y1, w_se)


def depth_warping(depth_1, depth_2, mask, translation, rotation, intrinsics, epsilon=1.0e-8):
    """`DepthWarpingLayer.forward` (models.py:460-465 -> `_depth_warping` :469-554).

    Returns (warped depth map 2 expressed in frame 1 [B,1,H,W], intersect mask in {0,1})."""
    b, _, h, w = depth_1.shape
    dtype = depth_1.dtype
    d1 = depth_1 * mask                                                         # :473
    d2 = depth_2 * mask                                                         # :474
    k_inv, w_vec, m_mat = _pose_terms(translation, rotation, intrinsics)
    x, y = _mesh(h, w, dtype, depth_1.device)
    m = m_mat.reshape(b, 3, 3, 1, 1)
    q = [(m[:, r, 0] * x[0] + m[:, r, 1] * y[0] + m[:, r, 2]).reshape(b, 1, h, w) for r in range(3)]
    wv = w_vec.reshape(b, 3, 1, 1, 1)
    eps = torch.full((), epsilon, dtype=dtype)
    z2 = wv[:, 2] + d1 * q[2]                                                    # :504-507
    z2 = torch.where(mask > 0.5, z2, eps)                                        # :509
    z2 = torch.where(z2 > 0.0, z2, eps)                                          # :510
    u2 = (wv[:, 0] + d1 * q[0]) / z2                                             # :513-520
    v2 = (wv[:, 1] + d1 * q[1]) / z2                                             # :522-529
    w2 = torch.bmm(intrinsics, translation.reshape(b, 3, 1)).reshape(b, 3)       # :531
    m2 = torch.bmm(torch.bmm(intrinsics, rotation), k_inv)                       # :532
    temp = (m2[:, 2, 0].reshape(b, 1, 1, 1) * x + m2[:, 2, 1].reshape(b, 1, 1, 1) * y
            + m2[:, 2, 2].reshape(b, 1, 1, 1))                                   # :534-538
    src = mask * (w2[:, 2].reshape(b, 1, 1, 1) + d2 * temp)                      # :539-541
    warped = bilinear_zero_pad(src, u2, v2)                                      # :546
    sampled_mask = bilinear_zero_pad(mask, u2, v2)
    intersect = torch.where(sampled_mask * mask >= 0.9, torch.ones((), dtype=dtype),
                            torch.zeros((), dtype=dtype))                       # :550-552
    return warped, intersect.detach()
