"""Oracle restatement of `utils.point_cloud_from_depth` (TEST INFRASTRUCTURE ONLY): /root/reference/utils.py:825-852,
the pure-Python double loop of the evaluation export (evaluate.py:337-341), vectorised with numpy in float32 -- the
arithmetic the loop performs on numpy float32 scalars under NumPy >= 2 promotion rules ((w - cx) / fx * z, each step
rounded to float32).  Pinned by tests/golden/export_a.npz, produced by executing the reference function's own source."""
import numpy as np


def point_cloud_from_depth(depth_map, color_img, mask_img, intrinsic_matrix, point_cloud_downsampling,
                           min_threshold=None, max_threshold=None):
    height, width, _ = color_img.shape
    k = np.asarray(intrinsic_matrix, dtype=np.float32)
    f_x, c_x, f_y, c_y = k[0, 0], k[0, 2], k[1, 1], k[1, 2]                    # :830-833
    hh, ww = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    keep = (hh % point_cloud_downsampling == 0) & (ww % point_cloud_downsampling == 0) & (mask_img > 0.5)   # :837
    b, g, r = color_img[..., 0], color_img[..., 1], color_img[..., 2]          # :842-844
    if max_threshold is not None and min_threshold is not None:                # :845-847
        keep &= (np.maximum(np.maximum(r, g), b) >= max_threshold) & (np.minimum(np.minimum(r, g), b) <= min_threshold)
    z = depth_map.astype(np.float32)                                           # :838
    x = ((ww.astype(np.float32) - c_x) / f_x) * z                              # :839
    y = ((hh.astype(np.float32) - c_y) / f_y) * z                              # :840
    cols = [x, y, z, r.astype(np.uint8).astype(np.float32), g.astype(np.uint8).astype(np.float32), b.astype(np.uint8).astype(np.float32)]
    return np.stack([c[keep] for c in cols], axis=1).astype(np.float32).reshape(-1, 6)     # :851-853 (row-major order)
