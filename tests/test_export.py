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
