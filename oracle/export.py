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


def get_torch_training_data(pair_extrinsics, pair_projections, pair_indexes, point_cloud, mask_boundary,
                            view_indexes_per_point, clean_point_list, visible_view_indexes):
    """Restatement of `utils.get_torch_training_data` (utils.py:460-612): per view, project the SfM points (float64),
    round to pixels (np.round: half to even), keep the visible inlier points that land inside the image, in front of the
    camera and on the boundary mask (== 255), and scatter depth / flow into sparse images (numpy's fancy assignment: with
    duplicate pixel indices the LAST point wins)."""
    h, w = mask_boundary.shape[0], mask_boundary.shape[1]
    pts = np.asarray(point_cloud).reshape((-1, 4))                                              # :476
    mask = mask_boundary.reshape(-1)
    uv, cam = [], []
    for i in range(2):                                                                          # :477-492
        p = np.einsum('ij,mj->mi', pair_projections[i], pts)
        uv.append(np.round(p / p[:, 2].reshape((-1, 1))))
        c = np.einsum('ij,mj->mi', pair_extrinsics[i], pts)
        cam.append(c / c[:, 3].reshape((-1, 1)))
    outs = {k: [] for k in ("dm", "d", "fm", "f")}
    for i in range(2):
        vis = np.asarray(view_indexes_per_point[:, visible_view_indexes.index(pair_indexes[i])]).reshape(-1)   # :501, 509
        ok = vis > 0.5
        if len(clean_point_list) != 0:                                                          # :503-506
            ok = ok & (np.asarray(clean_point_list).reshape(-1) > 0.5)
        ok = ok & (uv[i][:, 0] <= w - 1) & (uv[i][:, 0] >= 0) & (uv[i][:, 1] <= h - 1) & (uv[i][:, 1] >= 0) & (cam[i][:, 2] > 0)   # :521-525
        idx = np.where(ok)[0]
        loc = (uv[i][idx, 0] + uv[i][idx, 1] * w).astype(np.int32)                               # :527-529
        on_mask = mask[loc] == 255                                                              # :530-533
        idx, loc = idx[on_mask], loc[on_mask]
        flow = np.zeros((h * w, 2), dtype=np.float32)
        fmask = np.zeros((h * w, 1), dtype=np.float32)
        depth = np.zeros((h * w, 1), dtype=np.float32)
        dmask = np.zeros((h * w, 1), dtype=np.float32)
        fmask[loc, 0] = 1.0                                                                     # :534
        flow[loc, :] = uv[1 - i][idx, :2] - uv[i][idx, :2]                                      # :548-557
        flow[:, 0] /= w                                                                         # :559-562
        flow[:, 1] /= h
        bad = np.where((np.abs(flow[:, 0]) > 5.0) | (np.abs(flow[:, 1]) > 5.0))[0]              # :564-574
        fmask[bad, 0] = 0.0
        flow[bad, :] = 0.0
        depth[loc, 0] = cam[i][idx, 2]                                                          # :585-590
        dmask[loc, 0] = 1.0
        outs["dm"].append(dmask); outs["d"].append(depth); outs["fm"].append(fmask); outs["f"].append(flow)
    return (np.array(outs["dm"], dtype="float32").reshape((-1, h, w, 1)), np.array(outs["d"], dtype="float32").reshape((-1, h, w, 1)),
            np.array(outs["fm"], dtype="float32").reshape((-1, h, w, 1)), np.array(outs["f"], dtype="float32").reshape((-1, h, w, 2)))


def raster_scene(seed=91, h=64, w=80, m=700):
    """A synthetic SfM sequence for the rasteriser: two nearby cameras looking down +z, points in front of, behind and beside
    them, points sharing a view-1 ray (duplicate pixels with different depths: exercises last-point-wins), points almost
    on camera 2's principal plane (flow outliers), partial visibility and inlier flags."""
    rs = np.random.RandomState(seed)
    f = 0.9 * w
    k = np.array([[f, 0, w / 2.0 - 0.5], [0, f, h / 2.0 - 0.5], [0, 0, 1.0]])

    def pose(rx, ry, t):
        cx, sx, cy, sy = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry)
        r = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        e = np.eye(4)
        e[:3, :3] = r
        e[:3, 3] = t
        return e
    extr = [pose(0.02, -0.03, [0.01, 0.0, 0.02]), pose(-0.05, 0.08, [-0.12, 0.03, -0.45])]
    proj = [k @ e[:3, :] for e in extr]
    pts = np.concatenate([rs.uniform(-0.6, 0.6, (m, 2)), rs.uniform(-0.3, 1.6, (m, 1)), np.ones((m, 1))], axis=1)
    c1 = -extr[0][:3, :3].T @ extr[0][:3, 3]                    # camera centre 1: points c1 + s (x - c1) share a view-1 pixel
    for j in range(0, 120, 2):
        pts[m - 1 - j, :3] = c1 + rs.uniform(0.5, 1.5) * (pts[j + 200, :3] - c1)
    c2 = -extr[1][:3, :3].T @ extr[1][:3, 3]
    z2 = extr[1][2, :3]
    for j in range(300, 330):                                   # nearly on camera 2's principal plane -> huge view-2 pixel coordinates
        d = (pts[j, :3] - c2) @ z2
        pts[j, :3] -= (d - 1e-4 * rs.uniform(0.5, 2.0)) * z2
    n_views = 5
    vis = (rs.rand(m, n_views) > 0.25).astype(np.float32)
    clean = (rs.rand(m) > 0.1).astype(np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    mask = np.where(((yy - h / 2) / (0.48 * h)) ** 2 + ((xx - w / 2) / (0.47 * w)) ** 2 < 1.0, 255, 0).astype(np.uint8)
    return dict(extr=np.stack(extr), proj=np.stack(proj), pts=pts, vis=vis, clean=clean, mask=mask,
                visible_view_indexes=np.array([3, 7, 8, 12, 20]), pair_indexes=np.array([7, 12]))
