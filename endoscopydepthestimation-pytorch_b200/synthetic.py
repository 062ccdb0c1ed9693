"""Synthetic training batches with the shapes, value ranges and sparsity of the reference's
DataLoader output (`/root/reference/train.py:244-248`, `dataset.py:453-462`), following the
recipe in SURVEY.md section 8(d).  There is no dataset on the benchmark machine, so bench.py,
the tests and smoke() all draw their inputs from here.  Input generation only: nothing in this
file is on the measured path.
"""
import math
from typing import Dict

import torch

# camera_intrinsics_per_view of the reference's example sequence (fx = fy) at 1080x1920
_EXAMPLE_FOCAL = 677.171


def _rodrigues(axis_angle: torch.Tensor) -> torch.Tensor:
    """[B,3] axis-angle -> [B,3,3] rotation matrices."""
    theta = axis_angle.norm(dim=1, keepdim=True).clamp_min(1e-12)
    k = axis_angle / theta
    kx, ky, kz = k[:, 0], k[:, 1], k[:, 2]
    zero = torch.zeros_like(kx)
    kmat = torch.stack([zero, -kz, ky, kz, zero, -kx, -ky, kx, zero], dim=1).reshape(-1, 3, 3)
    eye = torch.eye(3, dtype=axis_angle.dtype).expand(axis_angle.shape[0], 3, 3)
    s = torch.sin(theta).reshape(-1, 1, 1)
    c = torch.cos(theta).reshape(-1, 1, 1)
    return eye + s * kmat + (1.0 - c) * (kmat @ kmat)


def _smooth_depth(b, h, w, gen, dtype):
    """Sum of three low-frequency sinusoids mapped to [0.3, 1.5] (depths normalised by the
    global scale in the reference, dataset.py:391-392)."""
    y = torch.linspace(0, 1, h, dtype=dtype).reshape(1, 1, h, 1)
    x = torch.linspace(0, 1, w, dtype=dtype).reshape(1, 1, 1, w)
    acc = torch.zeros(b, 1, h, w, dtype=dtype)
    for _ in range(3):
        fx = torch.rand(b, 1, 1, 1, generator=gen, dtype=dtype) * 2.0 + 0.5
        fy = torch.rand(b, 1, 1, 1, generator=gen, dtype=dtype) * 2.0 + 0.5
        ph = torch.rand(b, 1, 1, 1, generator=gen, dtype=dtype) * (2.0 * math.pi)
        acc = acc + torch.sin(2.0 * math.pi * (fx * x + fy * y) + ph)
    return 0.9 + 0.2 * acc          # in [0.3, 1.5]


def _flow_from_depth_plain(depth, mask, t, r, k):
    """Closed-form flow used only to fabricate the sparse SfM flow targets."""
    b, _, h, w = depth.shape
    k_inv = torch.linalg.inv(k.double())
    temp = k.double() @ r.double().transpose(1, 2)
    wv = (temp @ (-t.double())).reshape(b, 3, 1, 1)
    m = (temp @ k_inv).reshape(b, 3, 3, 1, 1)
    y = torch.arange(h, dtype=torch.float64).reshape(1, h, 1)
    x = torch.arange(w, dtype=torch.float64).reshape(1, 1, w)
    d = depth.double()[:, 0]
    q = [m[:, i, 0] * x + m[:, i, 1] * y + m[:, i, 2] for i in range(3)]
    z = wv[:, 2] + d * q[2]
    mk = mask.double()[:, 0]
    z = 1.0e30 * (1 - mk) + mk * z
    u = (wv[:, 0] + d * q[0]) / z
    v = (wv[:, 1] + d * q[1]) / z
    return torch.stack([(u - x) / w, (v - y) / h], dim=1).to(depth.dtype)


def make_batch(batch_size: int, height: int, width: int, seed: int = 10085, *, all_ones_boundary: bool = False,
               sparse_prob: float = 0.005, downsampling: float = None, dtype=torch.float32,
               device="cpu") -> Dict[str, torch.Tensor]:
    """One synthetic batch keyed like the names in train.py:244-248 (generated on the CPU from
    `seed` -- train.py:80 uses 10085 -- then moved to `device`)."""
    gen = torch.Generator().manual_seed(seed)
    b, h, w = batch_size, height, width
    if downsampling is None:
        downsampling = 1024.0 / h          # 4 at 256 rows, 2 at 512 (--input_downsampling, README.md:52)
    colors_1 = torch.rand(b, 3, h, w, generator=gen, dtype=dtype) * 2 - 1   # dataset.py:148 normalises to [-1,1]
    colors_2 = torch.rand(b, 3, h, w, generator=gen, dtype=dtype) * 2 - 1
    if all_ones_boundary:
        boundary = torch.ones(1, 1, h, w, dtype=dtype)
    else:
        yy = torch.arange(h, dtype=dtype).reshape(h, 1) - (h - 1) / 2.0
        xx = torch.arange(w, dtype=dtype).reshape(1, w) - (w - 1) / 2.0
        radius = 0.48 * min(h, w) * math.sqrt(2.0)
        boundary = ((yy * yy + xx * xx) <= radius * radius).to(dtype).reshape(1, 1, h, w)
    boundaries = boundary.expand(b, 1, h, w).contiguous()

    depth_gt_1 = _smooth_depth(b, h, w, gen, dtype)
    depth_gt_2 = _smooth_depth(b, h, w, gen, dtype)

    def sparse(prob):
        return (torch.rand(b, 1, h, w, generator=gen, dtype=dtype) < prob).to(dtype) * boundaries

    sdm_1, sdm_2 = sparse(sparse_prob), sparse(sparse_prob)
    # the reference resamples pairs with empty masks (dataset.py:372-375): guarantee >= 1 point
    for msk in (sdm_1, sdm_2):
        msk[:, 0, h // 2, w // 2] = 1.0
    sd_1, sd_2 = sdm_1 * depth_gt_1, sdm_2 * depth_gt_2

    axis = torch.randn(b, 3, generator=gen, dtype=dtype)
    axis = axis / axis.norm(dim=1, keepdim=True)
    angle = torch.rand(b, 1, generator=gen, dtype=dtype) * 0.1
    rot_1_wrt_2 = _rodrigues(axis * angle)
    tdir = torch.randn(b, 3, generator=gen, dtype=dtype)
    tdir = tdir / tdir.norm(dim=1, keepdim=True)
    tlen = torch.rand(b, 1, generator=gen, dtype=dtype) * 0.08 + 0.02
    trans_1_wrt_2 = (tdir * tlen).reshape(b, 3, 1)
    rot_2_wrt_1 = rot_1_wrt_2.transpose(1, 2).contiguous()                    # dataset.py:398-399
    trans_2_wrt_1 = -(rot_2_wrt_1 @ trans_1_wrt_2)

    focal = _EXAMPLE_FOCAL / downsampling
    intr = torch.tensor([[focal, 0.0, w / 2.0], [0.0, focal, h / 2.0], [0.0, 0.0, 1.0]], dtype=dtype)
    intrinsics = intr.reshape(1, 3, 3).expand(b, 3, 3).contiguous()

    def sparse_flow(depth_gt, flow_mask, t, r):
        flow = _flow_from_depth_plain(depth_gt, boundaries, t, r, intrinsics) * flow_mask
        bad = (flow.abs() > 5.0).any(dim=1, keepdim=True)                     # utils.py:567-574
        flow = torch.where(bad, torch.zeros((), dtype=dtype), flow)
        return flow, torch.where(bad, torch.zeros((), dtype=dtype), flow_mask)

    sfm_1, sfm_2 = sparse(sparse_prob), sparse(sparse_prob)
    sf_1, sfm_1 = sparse_flow(depth_gt_1, sfm_1, trans_1_wrt_2, rot_1_wrt_2)
    sf_2, sfm_2 = sparse_flow(depth_gt_2, sfm_2, trans_2_wrt_1, rot_2_wrt_1)

    batch = dict(colors_1=colors_1, colors_2=colors_2, sparse_depths_1=sd_1, sparse_depths_2=sd_2,
                 sparse_depth_masks_1=sdm_1, sparse_depth_masks_2=sdm_2, sparse_flows_1=sf_1, sparse_flows_2=sf_2,
                 sparse_flow_masks_1=sfm_1, sparse_flow_masks_2=sfm_2, boundaries=boundaries,
                 rotations_1_wrt_2=rot_1_wrt_2, rotations_2_wrt_1=rot_2_wrt_1,
                 translations_1_wrt_2=trans_1_wrt_2, translations_2_wrt_1=trans_2_wrt_1, intrinsics=intrinsics,
                 depth_gt_1=depth_gt_1, depth_gt_2=depth_gt_2)
    return {k: v.to(device) for k, v in batch.items()}


def jitter_depths(batch: Dict[str, torch.Tensor], seed: int = 1):
    """d1, d2 = d* x U(0.8, 1.2): stand-ins for network outputs in kernel-only benches."""
    gen = torch.Generator().manual_seed(seed)
    out = []
    for key in ("depth_gt_1", "depth_gt_2"):
        d = batch[key]
        noise = torch.rand(d.shape, generator=gen, dtype=d.dtype) * 0.4 + 0.8
        out.append(d * noise.to(d.device))
    return out


BATCH_KEYS_H2D = ("colors_1", "colors_2", "sparse_depths_1", "sparse_depths_2", "sparse_depth_masks_1",
                  "sparse_depth_masks_2", "sparse_flows_1", "sparse_flows_2", "sparse_flow_masks_1",
                  "sparse_flow_masks_2", "boundaries", "rotations_1_wrt_2", "rotations_2_wrt_1",
                  "translations_1_wrt_2", "translations_2_wrt_1", "intrinsics")   # train.py:255-270
