"""Coefficient tables of ``PIL.Image.resize(size, Image.BILINEAR)`` on 8-bit images, for the device kernel
``cvx_split_patches_u8`` (csrc/preprocess.cu).

The reference prepares every colposcopic image with ``img.resize((1024, 1024), Image.BILINEAR)`` on the uint8 image
(MultiModal Prediction/Graph_Structure(data_augmentation).py:151-161).  Pillow's resampler is not ``F.interpolate``:

  * it is separable and runs the horizontal pass first, ROUNDING TO UINT8 between the passes and at the end;
  * the triangle filter's support grows with the scale factor when the image is reduced (antialiasing), and the weights
    of every output pixel are renormalised over the taps that fall inside the image;
  * 8-bit images are filtered in fixed point: weights are rounded to 22 fractional bits, the accumulator starts at one
    half and is shifted down and clipped to [0, 255].

This module restates that published algorithm (Pillow ``src/libImaging/Resample.c``: ``precompute_coeffs``,
``normalize_coeffs_8bpc``, ``ImagingResampleHorizontal_8bpc`` / ``Vertical_8bpc``) as integer tables per axis; the kernel
then reproduces Pillow's arithmetic exactly, so the patches entering the encoder equal the reference's bit for bit
(tests/test_patch_encoder.py checks the kernel against the installed Pillow for enlarging and reducing sources).
"""
from __future__ import annotations

import functools
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _triangle(x):
    x = np.abs(x)
    return np.where(x < 1.0, 1.0 - x, 0.0)


def _keys_cubic(x):
    """Pillow's bicubic kernel (a = -0.5), support 2."""
    a = -0.5
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0,
                    np.where(x < 2.0, (((x - 5.0) * x + 8.0) * x - 4.0) * a, 0.0))


_KERNELS = {"bilinear": (_triangle, 1.0), "bicubic": (_keys_cubic, 2.0)}


def bilinear_tables(in_size: int, out_size: int):
    return resample_tables(in_size, out_size, "bilinear")


@functools.lru_cache(maxsize=512)
def resample_tables(in_size: int, out_size: int, kind: str = "bilinear"):
    """Per output coordinate: first source index, number of taps, fixed-point weights (``[out_size, ksize]`` int32) of
    ``Image.resize`` with the BILINEAR or BICUBIC filter."""
    kernel, base_support = _KERNELS[kind]
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = base_support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    # all output coordinates at once; every floating-point operation keeps the order of the C loop
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    lo = np.maximum(np.trunc(center - support + 0.5), 0.0).astype(np.int64)
    hi = np.minimum(np.trunc(center + support + 0.5), float(in_size)).astype(np.int64)
    n = hi - lo
    idx = np.arange(ksize, dtype=np.int64)[None, :]
    live = idx < n[:, None]
    x = (((idx + lo[:, None]).astype(np.float64) - center[:, None]) + 0.5) * ss
    w = np.where(live, kernel(x), 0.0)
    ww = np.cumsum(w, axis=1)[:, -1]                  # Pillow accumulates the normaliser left to right (+0.0 is exact)
    w = np.where((ww != 0.0)[:, None], w / np.where(ww != 0.0, ww, 1.0)[:, None], w)
    fixed = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS)).astype(np.int64)
    kk = np.where(live, fixed, 0).astype(np.int32)
    return lo.astype(np.int32), n.astype(np.int32), kk


def resize_u8_reference(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Numpy restatement of the two fixed-point passes on an ``[H, W, C]`` uint8 array (host-side check of the tables)."""
    h, w, _ = img.shape
    cur = img.astype(np.int64)
    if w != out_w:
        xmin, cnt, kk = bilinear_tables(w, out_w)
        out = np.empty((h, out_w, img.shape[2]), dtype=np.int64)
        for xx in range(out_w):
            seg = cur[:, xmin[xx]:xmin[xx] + cnt[xx], :]
            acc = (seg * kk[xx, :cnt[xx]][None, :, None].astype(np.int64)).sum(1) + (1 << (PRECISION_BITS - 1))
            out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
        cur = out
    if h != out_h:
        ymin, cnt, kk = bilinear_tables(h, out_h)
        out = np.empty((out_h, cur.shape[1], img.shape[2]), dtype=np.int64)
        for yy in range(out_h):
            seg = cur[ymin[yy]:ymin[yy] + cnt[yy], :, :]
            acc = (seg * kk[yy, :cnt[yy]][:, None, None].astype(np.int64)).sum(0) + (1 << (PRECISION_BITS - 1))
            out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255)
        cur = out
    return cur.astype(np.uint8)
