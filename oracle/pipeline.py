"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the image side of the input pipeline, "next" row N3 --
`utils.get_pair_color_imgs` (utils.py:441-457: cv2.imread -> cv2.resize(fx = fy = 1/downsampling) -> crop -> BGR2RGB) and the
normalisation `dataset.py:148, 446-447` applies to it (albumentations Normalize(mean 0.5, std 0.5, max 255) + img_to_tensor).

The arithmetic lives in two third-party dependencies of the reference:
  * OpenCV (unpinned by the reference; 4.13.0 in this image): `cv::resize`, INTER_LINEAR, 8-bit: separable bilinear filter in
    fixed point (modules/imgproc/src/resize.cpp: 11-bit coefficients `saturate_cast<short>(w * 2048)`, horizontal pass in int,
    vertical pass `(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2`; horizontal positions clamp the fractional weight
    at the borders, vertical ones clamp only the row index).  PINNED: `resize_linear_8u` is compared bit for bit with cv2.resize
    itself (tests/test_pipeline.py, and at fixture generation by executing the unmodified reference function).  For a factor of
    exactly 2 cv::resize switches to its INTER_AREA fast path, which agrees with this formula except on the last row / column of
    odd-sized sources: that case is rejected.
  * albumentations (absent from this image, unpinned): `Normalize` = `(float32(img) - mean * max) * reciprocal(std * max)` in
    float32 -- restated from its published source, PARITY UNPINNED for this one function (nothing here can execute it).
"""
import numpy as np


def _x_coefficients(dn, sn, scale):
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    frac = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    frac[lo] = 0
    s[lo] = 0
    hi = s >= sn - 1
    frac[hi] = 0
    s[hi] = sn - 1
    a1 = np.rint(frac * np.float32(2048)).astype(np.int64)
    a0 = np.rint((np.float32(1.0) - frac) * np.float32(2048)).astype(np.int64)
    return s, np.minimum(s + 1, sn - 1), a0, a1


def _y_coefficients(dn, sn, scale):
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    frac = (f - s.astype(np.float32)).astype(np.float32)
    b1 = np.rint(frac * np.float32(2048)).astype(np.int64)
    b0 = np.rint((np.float32(1.0) - frac) * np.float32(2048)).astype(np.int64)
    return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), b0, b1


def resized_shape(sh, sw, fx, fy):
    """dsize of cv2.resize(img, (0, 0), fx, fy): saturate_cast<int> = round half to even"""
    return int(np.rint(sh * fy)), int(np.rint(sw * fx))


def is_area_fast_2x(fx, fy):
    return abs(1.0 / fx - 2.0) < np.finfo(np.float64).eps and abs(1.0 / fy - 2.0) < np.finfo(np.float64).eps


def resize_linear_8u(img, fx, fy):
    """cv2.resize(img, (0, 0), fx=fx, fy=fy) for uint8 HxWxC images, bit for bit (see the module docstring)."""
    sh, sw = img.shape[:2]
    if is_area_fast_2x(fx, fy) and (sh % 2 or sw % 2):
        raise NotImplementedError("factor 2 on an odd-sized image: cv2.resize takes its INTER_AREA border path")
    dh, dw = resized_shape(sh, sw, fx, fy)
    x0, x1, a0, a1 = _x_coefficients(dw, sw, 1.0 / fx)
    y0, y1, b0, b1 = _y_coefficients(dh, sh, 1.0 / fy)
    src = img.astype(np.int64)
    rows = src[:, x0, :] * a0[None, :, None] + src[:, x1, :] * a1[None, :, None]
    s0, s1 = rows[y0], rows[y1]
    return ((((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2).astype(np.uint8)


def get_pair_color_imgs(decoded_bgr, start_h, end_h, start_w, end_w, downsampling_factor, is_hsv, rgb_mode):
    """`utils.get_pair_color_imgs` (utils.py:441-457) from the DECODED images on (the JPEG decode, cv2.imread, is host IO)."""
    if is_hsv:
        raise NotImplementedError("HSV mode (cv2.COLOR_BGR2HSV_FULL) is not restated")
    out = []
    for img in decoded_bgr:
        small = resize_linear_8u(img, 1.0 / downsampling_factor, 1.0 / downsampling_factor)[start_h:end_h, start_w:end_w, :]   # :446-447
        out.append(small[:, :, ::-1] if rgb_mode == "rgb" else small)                                                         # :451-452
    h, w, c = out[0].shape
    return np.asarray(out, dtype=np.uint8).reshape((-1, h, w, c))


def normalize_to_tensor(img_u8):
    """albumentations Normalize(mean 0.5, std 0.5, max_pixel_value 255) + img_to_tensor (dataset.py:148, 446-453): HxWx3 uint8 ->
    3xHxW float32."""
    mean = np.float32(0.5) * np.float32(255.0)
    den = np.reciprocal(np.float32(0.5) * np.float32(255.0), dtype=np.float32)
    x = img_u8.astype(np.float32)
    x -= mean
    x *= den
    return np.ascontiguousarray(np.moveaxis(x, -1, 0))
