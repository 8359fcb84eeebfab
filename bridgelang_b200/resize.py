"""resize.py — coefficient tables of Pillow's antialiased resize (ImagingResample, 8 bits per channel), built with the
same double-precision arithmetic Pillow uses, for the device kernel `blb_resize_u8` (csrc/resize.cu).

Restates src/libImaging/Resample.c of Pillow (`precompute_coeffs` + `normalize_coeffs_8bpc`, bicubic a = -0.5,
support 2) — the routine `PIL.Image.resize(size, BICUBIC)` reaches, i.e. what torchvision's `Resize(...,
interpolation=BICUBIC)` runs inside the reference's image transform (prismatic/models/backbones/vision/
dinosiglip_vit.py:91-111, prismatic/extern/hf/processing_prismatic.py:128-145).  Pillow is a third-party dependency of
the reference, not part of its tree; parity is pinned by comparing against PIL.Image.resize itself
(tests/test_resize.py) — bit for bit.
"""

from __future__ import annotations

import math
from functools import lru_cache
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _bilinear(x: float) -> float:
    if x < 0.0:
        x = -x
    return 1.0 - x if x < 1.0 else 0.0


_FILTERS = {"bicubic": (_bicubic, 2.0), "bilinear": (_bilinear, 1.0)}


@lru_cache(maxsize=64)
def resample_coeffs(in_size: int, out_size: int, interpolation: str = "bicubic") -> Tuple[np.ndarray, np.ndarray, int]:
    """(kk int32 [out_size, ksize], bounds int32 [out_size, 2] = (first input index, count), ksize) for one axis."""
    filt, fsupport = _FILTERS[interpolation]
    in0, in1 = 0.0, float(in_size)
    scale = filterscale = (in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:                                 # same summation order as the C loop
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return kk, bounds, ksize


def resize_u8_reference(frames: np.ndarray, out_hw: Tuple[int, int], interpolation: str = "bicubic") -> np.ndarray:
    """NumPy statement of the two fixed-point passes (uint8 [B,H,W,3] → [B,Hd,Wd,3]); host-side mirror of the kernel,
    used by the CPU tests to pin the tables against PIL."""
    B, Hs, Ws, C = frames.shape
    Hd, Wd = out_hw
    x = frames
    if Wd != Ws:
        kk, bounds, _ = resample_coeffs(Ws, Wd, interpolation)
        out = np.empty((B, Hs, Wd, C), dtype=np.uint8)
        for xo in range(Wd):
            x0, n = bounds[xo]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(x[:, :, x0:x0 + n, :].astype(np.int64), kk[xo, :n].astype(np.int64), axes=([2], [0]))
            out[:, :, xo, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
        x = out
    if Hd != Hs:
        kk, bounds, _ = resample_coeffs(Hs, Hd, interpolation)
        out = np.empty((B, Hd, x.shape[2], C), dtype=np.uint8)
        for yo in range(Hd):
            y0, n = bounds[yo]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(x[:, y0:y0 + n].astype(np.int64), kk[yo, :n].astype(np.int64), axes=([1], [0]))
            out[:, yo] = np.clip(acc >> PRECISION_BITS, 0, 255)
        x = out
    return x
