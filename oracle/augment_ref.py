"""TEST INFRASTRUCTURE - CPU restatement (numpy, integer arithmetic) of the reference's training augmentation
``DeeplabDataset.get_random_data`` (Segmentation/deeplabv3+/utils/dataloader.py:55-154), stage by stage.  Only ``tests/``
and ``__graft_entry__.smoke()`` may import this module; the product path is ``cervix_b200.utils.dataloader`` (CUDA).

The reference composes third-party primitives (Pillow 8-bit resampling, OpenCV 8-bit filters); those libraries are not
part of /root/reference, so each stage restates the library's published algorithm and is PINNED against the installed
library (Pillow 12.2, opencv-python 4.13) by tests/test_augment.py, and the composition is pinned against the
reference's own function run under fixed numpy seeds (tests/golden/augment_*.npz, made by oracle/make_golden_augment.py).

  stage                      reference line                 library algorithm restated here
  bicubic resize (image)     dataloader.py:91               Pillow Resample.c: precompute_coeffs / normalize_coeffs_8bpc /
                                                            ImagingResampleHorizontal_8bpc / Vertical_8bpc (22-bit fixed
                                                            point, horizontal pass first, uint8 between the passes)
  nearest resize (label)     dataloader.py:92               Pillow Geometry.c: ImagingScaleAffine (running double sum)
  flip, paste on the canvas  dataloader.py:97-113           pure indexing
  Gaussian blur 5x5          dataloader.py:120-122          OpenCV smooth: kernel (1,4,6,4,1)/16 for sigma = 0, 8.8 fixed
                                                            point per pass, BORDER_REFLECT_101
  rotation                   dataloader.py:127-133          OpenCV imgwarp: warpAffine coordinates in 10-bit fixed point,
                                                            5-bit sub-pixel phase, 15-bit bicubic table (A = -0.75),
                                                            constant border
  HSV jitter                 dataloader.py:139-153          OpenCV color_hsv: RGB2HSV_b (12-bit division tables),
                                                            HSV2RGB_f on floats scaled back to 8 bits
"""
from __future__ import annotations

import functools
import math

import numpy as np

# ------------------------------------------------------------------------------------------------ Pillow resampling
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0,
                    np.where(x < 2.0, (((x - 5.0) * x + 8.0) * x - 4.0) * a, 0.0))


def _bilinear(x):
    x = np.abs(x)
    return np.where(x < 1.0, 1.0 - x, 0.0)


_FILTERS = {"bicubic": (_bicubic, 2.0), "bilinear": (_bilinear, 1.0)}


@functools.lru_cache(maxsize=256)
def pil_coeffs(in_size: int, out_size: int, kind: str = "bicubic"):
    """Per output coordinate: first source index, tap count, 22-bit fixed-point weights."""
    filt, support = _FILTERS[kind]
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, dtype=np.int64)
    cnt = np.zeros(out_size, dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        n = hi - lo
        w = filt((np.arange(n, dtype=np.float64) + lo - center + 0.5) * ss)
        ww = 0.0
        for v in w:                                          # sequential left-to-right sum, as the C loop
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        fixed = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS)).astype(np.int64)
        xmin[xx], cnt[xx] = lo, n
        kk[xx, :n] = fixed
    return xmin, cnt, kk


def pil_resize_u8(img: np.ndarray, out_w: int, out_h: int, kind: str = "bicubic") -> np.ndarray:
    """``Image.fromarray(img).resize((out_w, out_h), BICUBIC)`` for an [H, W, C] uint8 array."""
    h, w = img.shape[:2]
    cur = img.astype(np.int64)
    if cur.ndim == 2:
        cur = cur[:, :, None]
    if w != out_w:
        xmin, cnt, kk = pil_coeffs(w, out_w, kind)
        out = np.empty((h, out_w, cur.shape[2]), dtype=np.int64)
        for xx in range(out_w):
            seg = cur[:, xmin[xx]:xmin[xx] + cnt[xx], :]
            acc = (seg * kk[xx, :cnt[xx]][None, :, None]).sum(1) + (1 << (PRECISION_BITS - 1))
            out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
        cur = out
    if h != out_h:
        ymin, cnt, kk = pil_coeffs(h, out_h, kind)
        out = np.empty((out_h, cur.shape[1], cur.shape[2]), dtype=np.int64)
        for yy in range(out_h):
            seg = cur[ymin[yy]:ymin[yy] + cnt[yy], :, :]
            acc = (seg * kk[yy, :cnt[yy]][:, None, None]).sum(0) + (1 << (PRECISION_BITS - 1))
            out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255)
        cur = out
    res = cur.astype(np.uint8)
    return res[:, :, 0] if img.ndim == 2 else res


def pil_nearest_index(in_size: int, out_size: int) -> np.ndarray:
    """Source index of every output coordinate of ``resize(..., NEAREST)``: the running double sum of Geometry.c."""
    a = float(in_size) / out_size
    idx = np.empty(out_size, dtype=np.int64)
    xo = 0.0 + a * 0.5
    for x in range(out_size):
        idx[x] = -1 if xo < 0.0 else int(xo)
        xo += a
    return np.clip(idx, 0, in_size - 1)


def pil_resize_nearest(lab: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    h, w = lab.shape[:2]
    if (w, h) == (out_w, out_h):
        return lab.copy()
    return lab[pil_nearest_index(h, out_h)][:, pil_nearest_index(w, out_w)]


def paste(canvas_hw, fill, src: np.ndarray, dx: int, dy: int) -> np.ndarray:
    """``Image.new(mode, (w, h), fill).paste(src, (dx, dy))`` (the part of src that falls on the canvas)."""
    h, w = canvas_hw
    out = np.empty((h, w) + src.shape[2:], dtype=np.uint8)
    out[...] = fill
    sh, sw = src.shape[:2]
    x0, y0 = max(dx, 0), max(dy, 0)
    x1, y1 = min(dx + sw, w), min(dy + sh, h)
    if x1 > x0 and y1 > y0:
        out[y0:y1, x0:x1] = src[y0 - dy:y1 - dy, x0 - dx:x1 - dx]
    return out


# ------------------------------------------------------------------------------------------------ OpenCV Gaussian blur
def _reflect101(i, n):
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * n - 2 - i, i)


def cv_gaussian5_u8(img: np.ndarray) -> np.ndarray:
    """``cv2.GaussianBlur(img, (5, 5), 0)`` on uint8: taps (1, 4, 6, 4, 1)/16 per pass in 8.8 fixed point; both passes
    are exact in 16 bits, so the result is round-half-up of the 25-tap integer sum / 256."""
    k = np.array([1, 4, 6, 4, 1], dtype=np.int64)
    h, w = img.shape[:2]
    src = img.astype(np.int64)
    tmp = np.zeros_like(src)
    cols = np.arange(w)
    for t in range(5):
        tmp += k[t] * src[:, _reflect101(cols + t - 2, w)]
    out = np.zeros_like(src)
    rows = np.arange(h)
    for t in range(5):
        out += k[t] * tmp[_reflect101(rows + t - 2, h)]
    return ((out + 128) >> 8).astype(np.uint8)


# ------------------------------------------------------------------------------------------------ OpenCV warpAffine
AB_BITS, INTER_BITS, COEF_BITS = 10, 5, 15
INTER_TAB = 1 << INTER_BITS


def _cv_round(x):
    """cvRound / saturate_cast<int>(double): round half to even."""
    return np.rint(x).astype(np.int64)


def rotation_matrix(w: int, h: int, rotation: int) -> np.ndarray:
    """``cv2.getRotationMatrix2D((w // 2, h // 2), -rotation, 1)`` followed by warpAffine's own inversion (doubles)."""
    cx, cy = float(w // 2), float(h // 2)
    ang = -rotation * (np.pi / 180.0)
    alpha, beta = math.cos(ang), math.sin(ang)
    m = [alpha, beta, (1 - alpha) * cx - beta * cy, -beta, alpha, beta * cx + (1 - alpha) * cy]
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0] = a11
    m[1] *= -d
    m[3] *= -d
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return np.array(m, dtype=np.float64)


def warp_coords(m: np.ndarray, w: int, h: int, nearest: bool):
    """Fixed-point source coordinates of every destination pixel: integer part and (cubic) 5-bit phases."""
    ab = float(1 << AB_BITS)
    xs = np.arange(w, dtype=np.float64)
    adelta = _cv_round(m[0] * xs * ab)
    bdelta = _cv_round(m[3] * xs * ab)
    ys = np.arange(h, dtype=np.float64)
    rd = (1 << AB_BITS) // 2 if nearest else (1 << AB_BITS) // INTER_TAB // 2
    x0 = _cv_round((m[1] * ys + m[2]) * ab) + rd
    y0 = _cv_round((m[4] * ys + m[5]) * ab) + rd
    if nearest:
        X = (x0[:, None] + adelta[None, :]) >> AB_BITS
        Y = (y0[:, None] + bdelta[None, :]) >> AB_BITS
        return X, Y, None, None
    X = (x0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    return X >> INTER_BITS, Y >> INTER_BITS, X & (INTER_TAB - 1), Y & (INTER_TAB - 1)


@functools.lru_cache(maxsize=1)
def cubic_table():
    """OpenCV's 32 x 32 table of 4 x 4 bicubic weights in 15-bit fixed point (initInterTab2D, fixpt)."""
    a = np.float32(-0.75)
    tab = np.zeros((INTER_TAB, 4), dtype=np.float32)
    scale = np.float32(1.0 / INTER_TAB)
    for i in range(INTER_TAB):
        x = np.float32(i) * scale
        one = np.float32(1)
        c0 = ((a * (x + one) - np.float32(5) * a) * (x + one) + np.float32(8) * a) * (x + one) - np.float32(4) * a
        c1 = ((a + np.float32(2)) * x - (a + np.float32(3))) * x * x + one
        xm = one - x
        c2 = ((a + np.float32(2)) * xm - (a + np.float32(3))) * xm * xm + one
        c3 = one - c0 - c1 - c2
        tab[i] = (c0, c1, c2, c3)
    out = np.zeros((INTER_TAB, INTER_TAB, 4, 4), dtype=np.int64)
    one15 = 1 << COEF_BITS
    for i in range(INTER_TAB):          # y phase
        for j in range(INTER_TAB):      # x phase
            it = np.zeros((4, 4), dtype=np.int64)
            for k1 in range(4):
                vy = tab[i, k1]
                for k2 in range(4):
                    v = np.float32(vy * tab[j, k2])
                    it[k1, k2] = int(np.clip(np.rint(np.float32(v * np.float32(one15))), -32768, 32767))
            isum = int(it.sum())
            if isum != one15:
                diff = isum - one15
                mk1 = mk2 = Mk1 = Mk2 = 2                  # OpenCV searches rows / columns ksize/2 .. ksize/2 + 1
                for k1 in range(2, 4):
                    for k2 in range(2, 4):
                        if it[k1, k2] < it[mk1, mk2]:
                            mk1, mk2 = k1, k2
                        elif it[k1, k2] > it[Mk1, Mk2]:
                            Mk1, Mk2 = k1, k2
                if diff < 0:
                    it[Mk1, Mk2] -= diff
                else:
                    it[mk1, mk2] -= diff
            out[i, j] = it
    return out


def cv_warp_cubic_u8(img: np.ndarray, rotation: int, border: int = 128) -> np.ndarray:
    """``cv2.warpAffine(img, getRotationMatrix2D(center, -rotation, 1), (w, h), flags=INTER_CUBIC, borderValue=border)``."""
    h, w = img.shape[:2]
    m = rotation_matrix(w, h, rotation)
    X, Y, fx, fy = warp_coords(m, w, h, nearest=False)
    wt = cubic_table()[fy, fx]                              # [h, w, 4, 4]
    src = img.astype(np.int64)
    acc = np.zeros((h, w, img.shape[2]), dtype=np.int64)
    for k1 in range(4):
        sy = Y - 1 + k1
        for k2 in range(4):
            sx = X - 1 + k2
            ok = (sx >= 0) & (sx < w) & (sy >= 0) & (sy < h)
            px = np.where(ok[..., None], src[np.clip(sy, 0, h - 1), np.clip(sx, 0, w - 1)], border)
            acc += px * wt[:, :, k1, k2][..., None]
    return np.clip((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS, 0, 255).astype(np.uint8)


def cv_warp_nearest_u8(lab: np.ndarray, rotation: int, border: int = 0) -> np.ndarray:
    h, w = lab.shape[:2]
    m = rotation_matrix(w, h, rotation)
    X, Y, _, _ = warp_coords(m, w, h, nearest=True)
    ok = (X >= 0) & (X < w) & (Y >= 0) & (Y < h)
    return np.where(ok, lab[np.clip(Y, 0, h - 1), np.clip(X, 0, w - 1)], border).astype(np.uint8)


# ------------------------------------------------------------------------------------------------ OpenCV HSV (8 bit)
HSV_SHIFT = 12


@functools.lru_cache(maxsize=1)
def hsv_div_tables():
    i = np.arange(1, 256, dtype=np.float64)
    sdiv = np.zeros(256, dtype=np.int64)
    hdiv = np.zeros(256, dtype=np.int64)
    sdiv[1:] = _cv_round((255 << HSV_SHIFT) / (1.0 * i))
    hdiv[1:] = _cv_round((180 << HSV_SHIFT) / (6.0 * i))
    return sdiv, hdiv


def cv_rgb2hsv_u8(img: np.ndarray) -> np.ndarray:
    """``cv2.cvtColor(img, COLOR_RGB2HSV)`` on uint8 (H in 0..179)."""
    sdiv, hdiv = hsv_div_tables()
    r, g, b = (img[..., c].astype(np.int64) for c in range(3))
    v = np.maximum(np.maximum(r, g), b)
    vmin = np.minimum(np.minimum(r, g), b)
    diff = v - vmin
    vr = v == r
    vg = v == g
    s = (diff * sdiv[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    hh = np.where(vr, g - b, np.where(vg, b - r + 2 * diff, r - g + 4 * diff))
    hh = (hh * hdiv[diff] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    hh = hh + np.where(hh < 0, 180, 0)
    return np.stack([hh, s, v], axis=-1).astype(np.uint8)


CV_SIMD_PIXELS = 32      # pixels per vector of OpenCV's AVX2 colour loops (opencv-python 4.13 on x86-64 with AVX2 + FMA3)


def _fnmadd32(a, b, c):
    """c - a*b rounded once to float32 (the FMA the AVX2 build contracts ``1 - s*h`` into)."""
    return (c.astype(np.float64) - a.astype(np.float64) * b.astype(np.float64)).astype(np.float32)


def cv_hsv2rgb_u8(hsv: np.ndarray) -> np.ndarray:
    """``cv2.cvtColor(hsv, COLOR_HSV2RGB)`` on uint8 [H, W, 3]: the float32 sector formula.  OpenCV runs every row in
    vectors of 32 pixels - there ``1 - s*h`` is one fused multiply-add and the result is TRUNCATED to 8 bits - and the
    remaining ``W % 32`` pixels of the row in scalar code: separate multiply and subtract, ROUNDED to nearest even.
    (Checked against cv2 for all 180 x 256 x 256 inputs and for ragged widths, tests/test_augment.py.)"""
    f = np.float32
    w = hsv.shape[1]
    vec = (np.arange(w) < (w // CV_SIMD_PIXELS) * CV_SIMD_PIXELS)[None, :]
    h = hsv[..., 0].astype(np.float32) * f(6.0 / 180.0)
    s = hsv[..., 1].astype(np.float32) * f(1.0 / 255.0)
    v = hsv[..., 2].astype(np.float32) * f(1.0 / 255.0)
    sector = np.trunc(h)
    frac = h - sector
    sector = sector.astype(np.int64) % 6
    one = np.ones_like(v)
    t1 = v * (one - s)
    t2 = np.where(vec, v * _fnmadd32(s, frac, one), v * (one - s * frac))
    t3 = np.where(vec, v * _fnmadd32(s, one - frac, one), v * (one - s * (one - frac)))
    tabs = np.stack([v, t1, t2, t3], axis=-1)
    sel = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]], dtype=np.int64)   # (b, g, r)
    bgr = np.take_along_axis(tabs, sel[sector], axis=-1)
    rgb = bgr[..., ::-1] * f(255.0)
    out = np.where(vec[..., None], np.trunc(rgb), np.rint(rgb))
    return np.clip(out, 0, 255).astype(np.uint8)


def hsv_luts(r):
    """The three 256-entry tables of dataloader.py:148-150 for gains r = (hue, sat, val)."""
    x = np.arange(0, 256, dtype=np.float64)
    return (((x * r[0]) % 180).astype(np.uint8), np.clip(x * r[1], 0, 255).astype(np.uint8),
            np.clip(x * r[2], 0, 255).astype(np.uint8))


def hsv_jitter_u8(img: np.ndarray, r) -> np.ndarray:
    hsv = cv_rgb2hsv_u8(img)
    lh, ls, lv = hsv_luts(r)
    hsv = np.stack([lh[hsv[..., 0]], ls[hsv[..., 1]], lv[hsv[..., 2]]], axis=-1)
    return cv_hsv2rgb_u8(hsv)


# ------------------------------------------------------------------------------------------------ the whole function
def draw_params(iw: int, ih: int, input_shape, rng=np.random, jitter=.3, hue=.1, sat=.7, val=.3):
    """The random decisions of get_random_data, drawn from ``rng`` in the reference's order (dataloader.py:81-139)."""
    h, w = input_shape
    rand = lambda a=0.0, b=1.0: rng.rand() * (b - a) + a   # noqa: E731
    new_ar = iw / ih * rand(1 - jitter, 1 + jitter) / rand(1 - jitter, 1 + jitter)
    scale = rand(0.25, 2)
    if new_ar < 1:
        nh = int(scale * h)
        nw = int(nh * new_ar)
    else:
        nw = int(scale * w)
        nh = int(nw / new_ar)
    flip = rand() < .5
    dx = int(rand(0, w - nw))
    dy = int(rand(0, h - nh))
    blur = rand() < 0.25
    rotate = rand() < 0.25
    rotation = int(rng.randint(-10, 11)) if rotate else 0
    r = rng.uniform(-1, 1, 3) * [hue, sat, val] + 1
    return dict(nw=nw, nh=nh, flip=bool(flip), dx=dx, dy=dy, blur=bool(blur), rotate=bool(rotate), rotation=rotation,
                r=np.asarray(r, dtype=np.float64))


def letterbox_params(iw: int, ih: int, input_shape):
    """The deterministic validation path (dataloader.py:64-77)."""
    h, w = input_shape
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    return dict(nw=nw, nh=nh, flip=False, dx=(w - nw) // 2, dy=(h - nh) // 2, blur=False, rotate=False, rotation=0, r=None)


def apply_params(image: np.ndarray, label: np.ndarray, input_shape, p):
    """get_random_data with its random decisions given: uint8 [ih, iw, 3] / [ih, iw] -> uint8 [h, w, 3] / [h, w]."""
    h, w = input_shape
    img = pil_resize_u8(image, p["nw"], p["nh"], "bicubic")
    lab = pil_resize_nearest(label, p["nw"], p["nh"])
    if p["flip"]:
        img, lab = img[:, ::-1], lab[:, ::-1]
    img = paste((h, w), 128, img, p["dx"], p["dy"])
    lab = paste((h, w), 0, lab, p["dx"], p["dy"])
    if p["blur"]:
        img = cv_gaussian5_u8(img)
    if p["rotate"]:
        img = cv_warp_cubic_u8(img, p["rotation"], 128)
        lab = cv_warp_nearest_u8(lab, p["rotation"], 0)
    if p["r"] is not None:
        img = hsv_jitter_u8(img, p["r"])
    return img, lab
