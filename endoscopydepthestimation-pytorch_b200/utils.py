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
