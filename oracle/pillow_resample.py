"""CPU oracle for the resize stage -- TEST INFRASTRUCTURE ONLY (see unet_oracle.py header).

``PIL.Image.resize((512, 512))`` (reference inference.py:35,63) is arithmetic of a third-party
dependency, Pillow (pinned ``Pillow==10.2.0``, reference requirements.txt:3; this image has 12.2).
This module restates its published algorithm (src/libImaging/Resample.c: ``bicubic_filter``,
``precompute_coeffs``, ``normalize_coeffs_8bpc``, ``ImagingResampleHorizontal_8bpc`` /
``Vertical_8bpc``) in numpy:

* kernel: bicubic, a = -0.5, support 2, stretched by max(scale, 1) (antialiasing);
* per output sample the taps are normalised in double precision, rounded to 22-bit fixed point;
* horizontal pass into a uint8 intermediate, then the vertical pass; each output is
  ``clip8((2**21 + sum(pixel * k)) >> 22)``.

Pinned by ``tests/test_prepost.py``: bit-identical to ``PIL.Image.resize`` executed in the test
process (Pillow is importable on the GPU box too), for up- and down-scaling and odd sizes.
"""
from __future__ import annotations

import math

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


def coeffs(in_size: int, out_size: int):
    """(kk int32 [out, ksize], bounds int32 [out, 2] = first tap, tap count) for one axis."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        for x, v in enumerate(k):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return kk, bounds


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_u8(img: np.ndarray, oh: int, ow: int) -> np.ndarray:
    """uint8 (H, W, C) -> uint8 (oh, ow, C), as ``Image.fromarray(img).resize((ow, oh))``."""
    h, w, c = img.shape
    cur = img
    if ow != w:
        kk, b = coeffs(w, ow)
        out = np.empty((h, ow, c), np.uint8)
        for xx in range(ow):
            x0, n = b[xx]
            acc = (cur[:, x0:x0 + n, :].astype(np.int64) * kk[xx, :n].astype(np.int64)[None, :, None]).sum(1)
            out[:, xx, :] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
        cur = out
    if oh != h:
        kk, b = coeffs(h, oh)
        out = np.empty((oh, cur.shape[1], c), np.uint8)
        for yy in range(oh):
            y0, n = b[yy]
            acc = (cur[y0:y0 + n].astype(np.int64) * kk[yy, :n].astype(np.int64)[:, None, None]).sum(0)
            out[yy] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
        cur = out
    return cur.copy() if cur is img else cur


def mask_bbox(mask: np.ndarray):
    """{xmin, xmax, ymin, ymax, count} of a boolean (H, W) mask, as inference.py:85-93 computes it
    with np.where; an empty mask gives (W, -1, H, -1, 0)."""
    ys, xs = np.where(mask)
    if len(xs) == 0:
        return (mask.shape[1], -1, mask.shape[0], -1, 0)
    return (int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max()), int(len(xs)))
