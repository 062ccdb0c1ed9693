"""Drop-in replacement for the export helper of /root/reference/utils.py that sits right after the network in
`evaluate.py:317-346`: `point_cloud_from_depth` (utils.py:825-852), a pure-Python H x W double loop in the reference,
one counting + one compacting CUDA kernel here (csrc/export.cu).  Same signature and return value (float32 [N, 6] numpy
array: x, y, z, r, g, b in row-major pixel order); inputs may be numpy arrays (uploaded) or CUDA tensors."""
import numpy as np
import torch

from . import _lib as L


def _to_cuda(a, dtype, device):
    t = torch.as_tensor(a)
    return t.to(device=device, dtype=dtype).contiguous()


def point_cloud_from_depth(depth_map, color_img, mask_img, intrinsic_matrix, point_cloud_downsampling,
                           min_threshold=None, max_threshold=None, device=None, return_tensor=False):
    if not torch.cuda.is_available():
        raise RuntimeError("endo_b200.utils.point_cloud_from_depth runs on a CUDA device only (there is no CPU fallback)")
    if device is None:
        device = depth_map.device if isinstance(depth_map, torch.Tensor) and depth_map.is_cuda else torch.device("cuda", torch.cuda.current_device())
    depth = _to_cuda(depth_map, torch.float32, device)
    color = _to_cuda(color_img, torch.uint8, device)
    mask = _to_cuda(mask_img, torch.float32, device)
    if color.dim() != 3 or color.shape[2] != 3:
        raise RuntimeError(f"color_img must be [H, W, 3], got {tuple(color.shape)}")
    h, w = color.shape[0], color.shape[1]
    if tuple(depth.shape) != (h, w) or tuple(mask.shape) != (h, w):
        raise RuntimeError(f"depth_map and mask_img must be [{h}, {w}], got {tuple(depth.shape)} and {tuple(mask.shape)}")
    k = np.asarray(intrinsic_matrix.detach().cpu() if isinstance(intrinsic_matrix, torch.Tensor) else intrinsic_matrix, dtype=np.float32)
    use_thr = max_threshold is not None and min_threshold is not None
    lib = L.lib()
    points = torch.empty((h * w, 6), dtype=torch.float32, device=device)
    count = torch.zeros(1, dtype=torch.int32, device=device)
    ws = torch.empty(lib.endo_point_cloud_workspace_bytes(h, w), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        L.check(lib.endo_point_cloud_from_depth(depth.data_ptr(), color.data_ptr(), mask.data_ptr(), float(k[0, 0]), float(k[1, 1]),
                                                float(k[0, 2]), float(k[1, 2]), h, w, int(point_cloud_downsampling), int(use_thr),
                                                float(min_threshold or 0.0), float(max_threshold or 0.0), points.data_ptr(),
                                                count.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr(device)),
                "point_cloud_from_depth")
    out = points[: int(count.item())]
    return out if return_tensor else out.cpu().numpy()


def get_torch_training_data(pair_extrinsics, pair_projections, pair_indexes, point_cloud, mask_boundary,
                            view_indexes_per_point, clean_point_list, visible_view_indexes, device=None, return_tensor=False):
    """`utils.get_torch_training_data` (utils.py:460-612) on the GPU: same arguments, same four return values
    (pair_depth_mask_imgs [2,H,W,1], pair_depth_imgs [2,H,W,1], pair_flow_mask_imgs [2,H,W,1], pair_flow_imgs [2,H,W,2],
    float32).  With `return_tensor=True` they stay on the device (the DataLoader's ten H2D image copies per pair,
    train.py:255-270, disappear)."""
    if not torch.cuda.is_available():
        raise RuntimeError("endo_b200.utils.get_torch_training_data runs on a CUDA device only (there is no CPU fallback)")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    mask_np = np.asarray(mask_boundary)
    h, w = mask_np.shape[0], mask_np.shape[1]
    pts = np.asarray(point_cloud, dtype=np.float64).reshape((-1, 4))
    m = pts.shape[0]
    vis = np.asarray(view_indexes_per_point)
    cols = [visible_view_indexes.index(pair_indexes[0]), visible_view_indexes.index(pair_indexes[1])]     # utils.py:501, 509
    vis2 = np.stack([np.asarray(vis[:, c]).reshape(-1) for c in cols]).astype(np.float32)
    clean = None if len(clean_point_list) == 0 else np.asarray(clean_point_list, dtype=np.float32).reshape(-1)
    proj = np.stack([np.asarray(p, dtype=np.float64).reshape(3, 4) for p in pair_projections[:2]])
    extr = np.stack([np.asarray(e, dtype=np.float64).reshape(4, 4) for e in pair_extrinsics[:2]])
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(device)
    d_pts, d_proj, d_extr, d_vis = t(pts, torch.float64), t(proj, torch.float64), t(extr, torch.float64), t(vis2, torch.float32)
    d_clean = None if clean is None else t(clean, torch.float32)
    d_mask = t(mask_np.reshape(h, w), torch.uint8)
    depth_mask = torch.empty((2, h, w, 1), dtype=torch.float32, device=device)
    depth = torch.empty((2, h, w, 1), dtype=torch.float32, device=device)
    flow_mask = torch.empty((2, h, w, 1), dtype=torch.float32, device=device)
    flow = torch.empty((2, h, w, 2), dtype=torch.float32, device=device)
    lib = L.lib()
    ws = torch.empty(lib.endo_rasterize_workspace_bytes(m, h, w), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        L.check(lib.endo_rasterize_pair(d_pts.data_ptr(), d_proj.data_ptr(), d_extr.data_ptr(), d_vis.data_ptr(), L.ptr(d_clean),
                                        d_mask.data_ptr(), m, h, w, depth_mask.data_ptr(), depth.data_ptr(), flow_mask.data_ptr(),
                                        flow.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr(device)), "rasterize_pair")
    out = (depth_mask, depth, flow_mask, flow)
    return out if return_tensor else tuple(o.cpu().numpy() for o in out)


def generating_pos_and_increment(idx, visible_view_indexes, adjacent_range):
    """`utils.generating_pos_and_increment` (utils.py:410-438), the pair sampler in front of the rasteriser: host logic, kept
    call-for-call compatible with the reference's use of the `random` module (same draws in the same order, so a seeded
    DataLoader worker yields the same pairs).  Returns [position in the visible-view list, signed offset of the partner]."""
    import random
    n = len(visible_view_indexes)
    pos = idx % n
    lo, hi = adjacent_range[0], adjacent_range[1]
    if n <= 2 * lo:                      # short sequence: shrink the minimum distance (utils.py:418-419)
        lo = n // 2
    forward_room, backward_room = min(hi, n - 1 - pos), min(hi, pos)
    if pos < lo:                         # too close to the start: the partner lies ahead
        return [pos, random.randint(lo, forward_room)]
    if pos >= n - lo:                    # too close to the end: the partner lies behind
        return [pos, -random.randint(lo, backward_room)]
    if random.randint(0, 1) == 1:        # interior: a coin decides the direction
        return [pos, random.randint(lo, forward_room)]
    return [pos, -random.randint(lo, backward_room)]


def resize_crop_color(img_bgr, start_h, end_h, start_w, end_w, downsampling_factor, rgb_mode="rgb", normalize=False, device=None):
    """One decoded frame (HxWx3 uint8, BGR as cv2.imread returns it; numpy or a CUDA tensor) -> cv2.resize(fx = fy =
    1/downsampling_factor) -> crop -> BGR2RGB (rgb_mode == "rgb"), bit for bit like utils.py:446-452, on the GPU.  Returns a
    CUDA uint8 tensor [H, W, 3]; with normalize=True also the float32 [3, H, W] tensor of dataset.py:148,446-453."""
    if not torch.cuda.is_available():
        raise RuntimeError("endo_b200.utils.resize_crop_color runs on a CUDA device only (there is no CPU fallback)")
    if device is None:
        device = img_bgr.device if isinstance(img_bgr, torch.Tensor) and img_bgr.is_cuda else torch.device("cuda", torch.cuda.current_device())
    src = img_bgr if isinstance(img_bgr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(img_bgr))
    if src.dtype != torch.uint8 or src.dim() != 3 or src.shape[2] != 3:
        raise RuntimeError("expected an HxWx3 uint8 image, got " + str(tuple(src.shape)) + " " + str(src.dtype))
    src = src.to(device).contiguous()
    h, w = int(end_h) - int(start_h), int(end_w) - int(start_w)
    if h <= 0 or w <= 0:
        raise RuntimeError("empty crop")
    out = torch.empty((h, w, 3), dtype=torch.uint8, device=device)
    norm = torch.empty((3, h, w), dtype=torch.float32, device=device) if normalize else None
    with torch.cuda.device(device):
        L.check(L.lib().endo_resize_crop_u8(src.data_ptr(), src.shape[0], src.shape[1], float(downsampling_factor), int(start_h), int(end_h),
                                            int(start_w), int(end_w), 1 if rgb_mode == "rgb" else 0, out.data_ptr(), L.ptr(norm),
                                            L.stream_ptr(device)), "resize_crop_u8")
    return (out, norm) if normalize else out


def get_pair_color_imgs(prefix_seq, pair_indexes, start_h, end_h, start_w, end_w, downsampling_factor, is_hsv, rgb_mode,
                        device=None, return_tensor=False):
    """`utils.get_pair_color_imgs` (utils.py:441-457), same arguments, same uint8 [2, H, W, 3] result: the JPEG decode stays
    cv2.imread on the host (no device decoder in this image), resize + crop + colour order run on the GPU."""
    if is_hsv:
        raise NotImplementedError("is_hsv=True (cv2.COLOR_BGR2HSV_FULL) is not implemented")
    import cv2
    from pathlib import Path
    imgs = []
    for i in pair_indexes:
        img = cv2.imread(str(Path(prefix_seq) / "{:08d}.jpg".format(i)))
        if img is None:
            raise RuntimeError("cannot read " + str(Path(prefix_seq) / "{:08d}.jpg".format(i)))
        imgs.append(resize_crop_color(img, start_h, end_h, start_w, end_w, downsampling_factor, rgb_mode, device=device))
    out = torch.stack(imgs)
    return out if return_tensor else out.cpu().numpy()
