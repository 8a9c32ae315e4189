"""TEST INFRASTRUCTURE — numpy restatement of the image preprocessing the reference delegates to third-party code:
Pillow's 8-bit bicubic resampler (Pillow is a dependency of the reference, requirements.txt; the algorithm restated is
src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal/Vertical_8bpc) as called by
data/transforms.py:15-76 (`img.resize(..., resample=BICUBIC)`), and ToTensor + Normalize (data/transforms.py:90-115).
Pinned against Pillow itself (present in this image and on the GPU box) by tests/test_host_misc.py."""
from __future__ import annotations

import math

import numpy as np


def bicubic_filter(x: float, a: float = -0.5) -> float:
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Scalar loop, operation for operation as Pillow: returns (int32 taps [out, ksize], int32 (xmin, n) [out, 2])."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 2.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / fscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        ww = 0.0
        for x in range(xmax):
            w = bicubic_filter((x + xmin - center + 0.5) * ss)
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    ik = np.where(kk < 0, np.trunc(-0.5 + kk * (1 << 22)), np.trunc(0.5 + kk * (1 << 22))).astype(np.int32)
    return ik, bounds


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    x = np.moveaxis(img, axis, 0).astype(np.int64)
    ik, b = precompute_coeffs(x.shape[0], out_size)
    out = np.empty((out_size,) + x.shape[1:], dtype=np.uint8)
    for o in range(out_size):
        xmin, n = b[o]
        acc = (1 << 21) + np.tensordot(ik[o, :n].astype(np.int64), x[xmin:xmin + n], axes=(0, 0))
        out[o] = np.clip(acc >> 22, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """uint8 [H, W, C] -> [out_h, out_w, C]: horizontal pass, then vertical pass on the uint8 intermediate."""
    h, w = img.shape[:2]
    x = img
    if out_w != w:
        x = _resample_axis(x, out_w, 1)
    if out_h != h:
        x = _resample_axis(x, out_h, 0)
    return x
