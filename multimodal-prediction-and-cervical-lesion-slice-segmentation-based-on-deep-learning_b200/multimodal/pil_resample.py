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


@functools.lru_cache(maxsize=64)
def bilinear_tables(in_size: int, out_size: int):
    """Per output coordinate: first source index, number of taps, fixed-point weights (``[out_size, ksize]`` int32)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale                       # bilinear: support 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, dtype=np.int32)
    cnt = np.zeros(out_size, dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        x = (np.arange(n, dtype=np.float64) + lo - center + 0.5) * ss
        w = np.where(np.abs(x) < 1.0, 1.0 - np.abs(x), 0.0)
        ww = float(w.sum())
        if ww != 0.0:
            w = w / ww
        fixed = np.where(w < 0, (-0.5 + w * (1 << PRECISION_BITS)), (0.5 + w * (1 << PRECISION_BITS))).astype(np.int64)
        xmin[xx], cnt[xx] = lo, n
        kk[xx, :n] = fixed.astype(np.int32)
    return xmin, cnt, kk


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
