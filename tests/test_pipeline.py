"""Image side of the input pipeline ("next" row N3): utils.get_pair_color_imgs (utils.py:441-457) after the JPEG decode, and the
normalisation of dataset.py:148,446-453.  CPU: the oracle restatement against cv2.resize itself and against the fixture recorded
by executing the unmodified reference function; GPU: the CUDA kernel bit for bit against both."""
import numpy as np
import pytest
import torch

import endo_b200
from oracle import pipeline as opipe
from conftest import load_golden

cv2 = pytest.importorskip("cv2")


def _cases(g):
    for ds, sh, eh, sw, ew, rgb in g["cases"].tolist():
        mode = "rgb" if rgb else "bgr"
        yield float(ds), int(sh), int(eh), int(sw), int(ew), mode, g[f"out_ds{ds}_{mode}"]


def test_oracle_resize_is_cv2_resize_bit_for_bit():
    rs = np.random.RandomState(1)
    for (h, w) in ((270, 480), (123, 77), (200, 333)):
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        for ds in (4.0, 3.0, 8.0, 2.5, 1.0, 1.7, 1.3, 5.0, 0.8):
            assert np.array_equal(opipe.resize_linear_8u(img, 1.0 / ds, 1.0 / ds), cv2.resize(img, (0, 0), fx=1.0 / ds, fy=1.0 / ds)), (h, w, ds)
    even = rs.randint(0, 256, (120, 96, 3)).astype(np.uint8)
    assert np.array_equal(opipe.resize_linear_8u(even, 0.5, 0.5), cv2.resize(even, (0, 0), fx=0.5, fy=0.5))
    with pytest.raises(NotImplementedError):
        opipe.resize_linear_8u(even[:119], 0.5, 0.5)


def test_oracle_matches_reference_fixture():
    g = load_golden("pipeline_a")
    frames = [g["decoded_0"], g["decoded_1"]]
    for k in range(2):          # the decoder of this container reproduces the frames the fixture was made from
        assert np.array_equal(cv2.imdecode(g[f"jpeg_{k}"], cv2.IMREAD_COLOR), frames[k])
    for ds, sh, eh, sw, ew, mode, want in _cases(g):
        got = opipe.get_pair_color_imgs(frames, sh, eh, sw, ew, ds, False, mode)
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), (ds, mode)
    x = opipe.normalize_to_tensor(g["out_ds4.0_rgb"][0])
    assert x.shape == (3, 48, 88) and x.dtype == np.float32 and float(x.min()) >= -1.0 and float(x.max()) <= 1.0


@pytest.mark.gpu
def test_cuda_resize_crop_is_bit_exact(tmp_path):
    g = load_golden("pipeline_a")
    frames = [g["decoded_0"], g["decoded_1"]]
    for k, idx in enumerate((17, 23)):
        (tmp_path / "{:08d}.jpg".format(idx)).write_bytes(g[f"jpeg_{k}"].tobytes())
    for ds, sh, eh, sw, ew, mode, want in _cases(g):
        got = endo_b200.utils.get_pair_color_imgs(str(tmp_path), [17, 23], sh, eh, sw, ew, ds, False, mode)   # the reference's call
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), (ds, mode)
        u8, norm = endo_b200.utils.resize_crop_color(frames[0], sh, eh, sw, ew, ds, mode, normalize=True)
        assert np.array_equal(u8.cpu().numpy(), want[0])
        assert np.array_equal(norm.cpu().numpy(), opipe.normalize_to_tensor(want[0]))
    # full-size frame, the reference's default factor, against cv2 directly
    rs = np.random.RandomState(5)
    big = rs.randint(0, 256, (1080, 1920, 3)).astype(np.uint8)
    ref = cv2.resize(big, (0, 0), fx=0.25, fy=0.25)[7:263, 80:400]
    got = endo_b200.utils.resize_crop_color(big, 7, 263, 80, 400, 4.0, "bgr")
    assert np.array_equal(got.cpu().numpy(), ref)
    with pytest.raises(RuntimeError):
        endo_b200.utils.resize_crop_color(big, 0, 300, 0, 100, 4.0)          # crop outside the 270 x 480 resized image
    with pytest.raises(NotImplementedError):
        endo_b200.utils.get_pair_color_imgs(str(tmp_path), [17, 23], 0, 8, 0, 8, 4.0, True, "rgb")
