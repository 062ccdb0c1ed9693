"""Evaluation export step after the network: utils.point_cloud_from_depth (utils.py:825-852).
CPU: the numpy oracle against the fixture produced by executing the reference function's own source; GPU: the CUDA
kernel against the oracle and the fixture, element for element (row-major order, fp32 operation order)."""
import numpy as np
import pytest
import torch

import endo_b200
from oracle import export as oexp
from conftest import load_golden


def _cases(g):
    return (("pc_all", dict(point_cloud_downsampling=1)), ("pc_ds2", dict(point_cloud_downsampling=2)),
            ("pc_thr", dict(point_cloud_downsampling=1, min_threshold=60, max_threshold=200)))


def test_oracle_matches_reference_fixture():
    g = load_golden("export_a")
    for key, kw in _cases(g):
        got = oexp.point_cloud_from_depth(g["depth"], g["color"], g["mask"], g["k"], **kw)
        assert got.shape == g[key].shape and got.dtype == np.float32
        assert np.array_equal(got, g[key]), key                    # bit-exact: same float32 operation order


@pytest.mark.gpu
def test_cuda_point_cloud_is_bit_exact():
    g = load_golden("export_a")
    for key, kw in _cases(g):
        got = endo_b200.utils.point_cloud_from_depth(g["depth"], g["color"], g["mask"], g["k"], **kw)
        assert got.shape == g[key].shape, (key, got.shape, g[key].shape)
        assert np.array_equal(got, g[key]), key
    # evaluate.py-sized image (256 x 320), CUDA tensors in, against the oracle
    rs = np.random.RandomState(5)
    h, w = 256, 320
    depth = (0.3 + rs.rand(h, w)).astype(np.float32)
    color = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    mask = (rs.rand(h, w) > 0.3).astype(np.float32)
    k = np.array([[169.3, 0, 160.0], [0, 169.3, 128.0], [0, 0, 1]], dtype=np.float32)
    ref = oexp.point_cloud_from_depth(depth, color, mask, k, 1)
    got = endo_b200.utils.point_cloud_from_depth(torch.tensor(depth).cuda(), torch.tensor(color).cuda(), torch.tensor(mask).cuda(), k, 1)
    assert np.array_equal(got, ref)
    with pytest.raises(RuntimeError):
        endo_b200.utils.point_cloud_from_depth(depth[:10], color, mask, k, 1)


# ------------------------------------------------------------------ sparse rasteriser (utils.get_torch_training_data, utils.py:460-612)
def _raster_args(g, clean):
    return ([g["extr"][0], g["extr"][1]], [g["proj"][0], g["proj"][1]], list(g["pair_indexes"]), [list(p) for p in g["pts"]],
            g["mask"], g["vis"], (g["clean"] if clean else []), list(g["visible_view_indexes"]))


RASTER_KEYS = ("depth_mask", "depth", "flow_mask", "flow")


def test_raster_fixture_exercises_duplicates_and_outliers():
    g = load_golden("raster_a")
    dm, fm = g["noclean_depth_mask"], g["noclean_flow_mask"]
    assert fm.sum() < dm.sum()                                     # |flow| > 5 outliers were cleared (utils.py:564-574)
    out = oexp.get_torch_training_data(*_raster_args(g, False))
    # several points land on one pixel: more valid points than drawn pixels
    uv = np.round((g["pts"] @ g["proj"][0].T) / (g["pts"] @ g["proj"][0].T)[:, 2:3])
    loc = uv[:, 0] + uv[:, 1] * g["mask"].shape[1]
    assert len(np.unique(loc)) < len(loc)
    assert out[0].shape == dm.shape


@pytest.mark.parametrize("clean", [True, False])
def test_raster_oracle_matches_reference_fixture(clean):
    g = load_golden("raster_a")
    out = oexp.get_torch_training_data(*_raster_args(g, clean))
    for key, o in zip(RASTER_KEYS, out):
        want = g[("clean_" if clean else "noclean_") + key]
        assert o.dtype == want.dtype and o.shape == want.shape
        assert np.array_equal(o, want), key


@pytest.mark.gpu
@pytest.mark.parametrize("clean", [True, False])
def test_cuda_rasteriser_is_bit_exact(clean):
    g = load_golden("raster_a")
    out = endo_b200.utils.get_torch_training_data(*_raster_args(g, clean))
    for key, o in zip(RASTER_KEYS, out):
        want = g[("clean_" if clean else "noclean_") + key]
        assert o.dtype == want.dtype and o.shape == want.shape
        assert np.array_equal(o, want), key
    dev = endo_b200.utils.get_torch_training_data(*_raster_args(g, clean), return_tensor=True)
    assert all(t.is_cuda for t in dev)


@pytest.mark.gpu
def test_cuda_rasteriser_empty_point_cloud_and_larger_image():
    g = load_golden("raster_a")
    a = list(_raster_args(g, False))
    a[3] = []
    a[5] = np.zeros((0, g["vis"].shape[1]), np.float32)
    out = endo_b200.utils.get_torch_training_data(*a)
    assert all(float(np.abs(o).sum()) == 0.0 for o in out)
    # a second, seeded 256 x 320 scene against the oracle
    sc = oexp.raster_scene(seed=5, h=256, w=320, m=3000)
    args = _raster_args(sc, True)
    want = oexp.get_torch_training_data(*args)
    got = endo_b200.utils.get_torch_training_data(*args)
    for key, o, wnt in zip(RASTER_KEYS, got, want):
        assert np.array_equal(o, wnt), key


# ------------------------------------------------------------------ pair sampler (utils.generating_pos_and_increment, utils.py:410-438)
def test_pair_sampler_draws_the_reference_pairs():
    """Host logic: with the same `random` seed the mirror returns what the unmodified reference function returned (fixture)."""
    import random
    g = load_golden("sampler_a")
    branches = set()
    for (n, lo, hi, idx, seed), (pos, inc) in zip(g["cases"].tolist(), g["out"].tolist()):
        views = list(range(100, 100 + 3 * n, 3))
        random.seed(1000 * seed + idx)
        got = endo_b200.utils.generating_pos_and_increment(idx, views, [lo, hi])
        assert got == [pos, inc], (n, lo, hi, idx, seed, got, pos, inc)
        assert 0 <= pos + inc < n
        branches.add("start" if pos < min(lo, n // 2 if n <= 2 * lo else lo) else ("end" if inc < 0 else "fwd"))
    assert len(g["cases"]) > 200 and len(branches) == 3
