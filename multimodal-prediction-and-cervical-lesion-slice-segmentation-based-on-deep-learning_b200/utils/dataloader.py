"""Segmentation loader with the training augmentation on the device (drop-in for the reference's ``utils/dataloader.py``).

Reference: ``DeeplabDataset`` / ``deeplab_dataset_collate``, Segmentation/deeplabv3+/utils/dataloader.py:12-169.  There a
DataLoader worker decodes one image, runs ``get_random_data`` on it with PIL and OpenCV (bicubic resize to a jittered
size, flip, paste on a grey canvas, Gaussian blur, rotation, HSV gain jitter), divides by 255, expands the class map to a
one-hot float array and ships 26 bytes per pixel to the GPU; four workers (train.py:281) deliver tens of images per
second.  Here a worker only DECODES and DRAWS the random decisions (from ``np.random`` in the reference's order, so a
seeded run makes the same decisions); the decoded uint8 sources of a batch cross PCIe once and four kernels
(csrc/augment.cu) do the arithmetic, bit for bit what Pillow 8-bit resampling and OpenCV's 8-bit filters produce.  The
result - uint8 ``[B,H,W,3]`` pixels and uint8 ``[B,H,W]`` class maps on the device - is what ``DeepLab.forward``,
``SegTrainer.step`` and ``fit_one_epoch`` take directly (the /255, the clamp of the ignore label and the one-hot target
happen inside the first kernel of the step and inside the loss kernel).

    ds  = DeeplabDataset(lines, input_shape, num_classes, True, dataset_path)
    gen = DeviceAugmentLoader(DataLoader(ds, batch_size=B, collate_fn=deeplab_dataset_collate, num_workers=4))
    fit_one_epoch(model_train, model, ..., gen=gen, ...)          # batches: (uint8 pixels, uint8 class maps, None)

Host-side tables (all integer, all cached): Pillow's coefficient rows per (source, target) size
(multimodal/pil_resample.py), Pillow's nearest-neighbour index rows, OpenCV's warpAffine fixed-point coordinate rows per
angle, OpenCV's 15-bit bicubic weight table, the three HSV gain tables of dataloader.py:148-150.
"""
from __future__ import annotations

import functools
import math
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ..multimodal.pil_resample import resample_tables

# cvx_aug_sample of include/cervix_b200.h (96 bytes)
AUG_SAMPLE = np.dtype([("src_off", "<i8"), ("lab_off", "<i8"), ("tmp_off", "<i8"), ("xtab", "<i4"), ("ytab", "<i4"),
                       ("xnn", "<i4"), ("ynn", "<i4"), ("rot", "<i4"), ("lut", "<i4"), ("ih", "<i4"), ("iw", "<i4"),
                       ("nh", "<i4"), ("nw", "<i4"), ("xtaps", "<i4"), ("ytaps", "<i4"), ("dx", "<i4"), ("dy", "<i4"),
                       ("flip", "<i4"), ("blur", "<i4"), ("rotate", "<i4"), ("reserved", "<i4")])
CV_VECTOR_PIXELS = 32        # pixels per vector of OpenCV's AVX2 colour loops; the ragged rest of a row runs scalar code


# ------------------------------------------------------------------------------------------------ random decisions
def draw_params(iw: int, ih: int, input_shape, rng=np.random, jitter=.3, hue=.1, sat=.7, val=.3) -> dict:
    """The random decisions of ``get_random_data`` (dataloader.py:81-139), drawn from ``rng`` in the reference's order:
    two aspect jitters, scale, flip, paste offsets, blur, rotate (+ angle), three HSV gains."""
    h, w = input_shape

    def rand(a=0.0, b=1.0):
        return rng.rand() * (b - a) + a

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


def letterbox_params(iw: int, ih: int, input_shape) -> dict:
    """The validation path (``random=False``, dataloader.py:64-77): aspect-preserving resize, centred on the canvas."""
    h, w = input_shape
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    return dict(nw=nw, nh=nh, flip=False, dx=(w - nw) // 2, dy=(h - nh) // 2, blur=False, rotate=False, rotation=0, r=None)


# ------------------------------------------------------------------------------------------------ integer tables
@functools.lru_cache(maxsize=512)
def nearest_table(in_size: int, out_size: int) -> np.ndarray:
    """Source index per output coordinate of ``Image.resize(..., NEAREST)``: Pillow steps a double by in/out from half a
    step and truncates (Geometry.c, affine scale) - the running sum, not a product, decides the ties."""
    a = float(in_size) / out_size
    steps = np.full(out_size, a, dtype=np.float64)
    steps[0] = a * 0.5
    pos = np.cumsum(steps)                           # sequential adds, as the C loop
    return np.clip(pos.astype(np.int64), 0, in_size - 1).astype(np.int32)


@functools.lru_cache(maxsize=64)
def rotation_tables(w: int, h: int, rotation: int) -> np.ndarray:
    """``adelta[w], bdelta[w], x0[h], y0[h]`` of ``cv2.warpAffine(img, getRotationMatrix2D((w//2, h//2), -rotation, 1))``:
    the inverse map in 10-bit fixed point, columns and rows rounded separately (imgwarp.cpp); the kernel adds the
    interpolation's rounding term."""
    cx, cy = float(w // 2), float(h // 2)
    ang = -rotation * (np.pi / 180.0)
    alpha, beta = math.cos(ang), math.sin(ang)
    m = [alpha, beta, (1 - alpha) * cx - beta * cy, -beta, alpha, beta * cx + (1 - alpha) * cy]
    d = m[0] * m[4] - m[1] * m[3]                    # warpAffine inverts the forward matrix itself
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0] = a11
    m[1] *= -d
    m[3] *= -d
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    ab = 1024.0
    xs, ys = np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64)
    parts = [np.rint(m[0] * xs * ab), np.rint(m[3] * xs * ab), np.rint((m[1] * ys + m[2]) * ab), np.rint((m[4] * ys + m[5]) * ab)]
    return np.concatenate(parts).astype(np.int32)


@functools.lru_cache(maxsize=1)
def cubic_weights() -> np.ndarray:
    """OpenCV's remap table for INTER_CUBIC: for each of 32 x 32 sub-pixel phases the 4 x 4 weights (A = -0.75, float32)
    scaled to 15 bits, with the rounding residue folded into the largest / smallest of the lower-right four so that every
    set sums to 32768 (imgwarp.cpp initInterTab2D).  int16 [32, 32, 16]."""
    f = np.float32
    a = f(-0.75)
    x = np.arange(32, dtype=np.float32) * f(1.0 / 32)
    one = f(1)
    c0 = ((a * (x + one) - f(5) * a) * (x + one) + f(8) * a) * (x + one) - f(4) * a
    c1 = ((a + f(2)) * x - (a + f(3))) * x * x + one
    xm = one - x
    c2 = ((a + f(2)) * xm - (a + f(3))) * xm * xm + one
    c3 = one - c0 - c1 - c2
    tab = np.stack([c0, c1, c2, c3], axis=1).astype(np.float32)                      # [32, 4]
    prod = (tab[:, None, :, None] * tab[None, :, None, :]).astype(np.float32)       # [y phase, x phase, k1, k2]
    it = np.clip(np.rint(prod * f(32768)), -32768, 32767).astype(np.int64)
    diff = it.sum(axis=(2, 3)) - 32768
    for i in range(32):
        for j in range(32):
            if diff[i, j] == 0:
                continue
            blk = it[i, j]
            mk = Mk = (2, 2)
            for k1 in (2, 3):
                for k2 in (2, 3):
                    if blk[k1, k2] < blk[mk]:
                        mk = (k1, k2)
                    elif blk[k1, k2] > blk[Mk]:
                        Mk = (k1, k2)
            blk[Mk if diff[i, j] < 0 else mk] -= diff[i, j]
    return it.reshape(32, 32, 16).astype(np.int16)


def hsv_luts(r) -> np.ndarray:
    """The three 256-entry tables of dataloader.py:147-150 for gains ``r = (hue, sat, val)``, as one uint8 [768] row."""
    x = np.arange(0, 256, dtype=np.float64)
    return np.concatenate([((x * r[0]) % 180).astype(np.uint8), np.clip(x * r[1], 0, 255).astype(np.uint8),
                           np.clip(x * r[2], 0, 255).astype(np.uint8)])


# ------------------------------------------------------------------------------------------------ packing a batch
class AugmentPlan:
    """One batch, packed for the device: descriptors, sources, integer tables, colour tables."""

    __slots__ = ("samples", "src", "tables", "luts", "shape", "max_elems", "tmp_bytes")

    def __init__(self, samples, src, tables, luts, shape, max_elems, tmp_bytes):
        self.samples, self.src, self.tables, self.luts = samples, src, tables, luts
        self.shape, self.max_elems, self.tmp_bytes = shape, max_elems, tmp_bytes

    def __len__(self):
        return len(self.samples)

    def pin_memory(self):
        for k in ("src", "tables", "luts"):
            setattr(self, k, getattr(self, k).pin_memory())
        return self


def _as_u8(a, ndim) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != ndim:
        raise TypeError("uint8 array with %d dimensions expected, got %s %s" % (ndim, a.dtype, a.shape))
    return np.ascontiguousarray(a)


def pack_batch(items: Sequence[Tuple[np.ndarray, np.ndarray, dict]], input_shape) -> AugmentPlan:
    """``items``: per image the decoded RGB pixels uint8 [ih, iw, 3], the class map uint8 [ih, iw] and the decisions of
    ``draw_params`` / ``letterbox_params``."""
    h, w = int(input_shape[0]), int(input_shape[1])
    n = len(items)
    samples = np.zeros(n, dtype=AUG_SAMPLE)
    src_parts: List[np.ndarray] = []
    tab_parts: List[np.ndarray] = []
    tab_index = {}
    luts = np.zeros((n, 768), dtype=np.uint8)
    src_off = tab_off = tmp_off = 0
    max_elems = 0

    def table(key, make):
        nonlocal tab_off
        if key not in tab_index:
            t = np.ascontiguousarray(make(), dtype=np.int32).reshape(-1)
            tab_index[key] = tab_off
            tab_parts.append(t)
            tab_off += t.size
        return tab_index[key]

    for i, (img, lab, p) in enumerate(items):
        img, lab = _as_u8(img, 3), _as_u8(lab, 2)
        ih, iw = lab.shape
        if img.shape != (ih, iw, 3):
            raise ValueError("image %s and class map %s differ in size" % (img.shape, lab.shape))
        nw, nh = int(p["nw"]), int(p["nh"])
        if nw <= 0 or nh <= 0:
            raise ValueError("height and width must be > 0")          # what PIL's resize raises for such a draw
        s = samples[i]
        s["ih"], s["iw"], s["nh"], s["nw"] = ih, iw, nh, nw
        s["dx"], s["dy"], s["flip"], s["blur"], s["rotate"] = p["dx"], p["dy"], p["flip"], p["blur"], p["rotate"]
        s["src_off"] = src_off
        src_parts.append(img.reshape(-1))
        src_off += img.size
        s["lab_off"] = src_off
        src_parts.append(lab.reshape(-1))
        src_off += lab.size
        pad = (-src_off) % 16
        if pad:
            src_parts.append(np.zeros(pad, dtype=np.uint8))
            src_off += pad
        xt, yt = resample_tables(iw, nw, "bicubic"), resample_tables(ih, nh, "bicubic")
        s["xtaps"], s["ytaps"] = xt[2].shape[1], yt[2].shape[1]
        s["xtab"] = table(("x", iw, nw), lambda: np.concatenate([xt[0], xt[1], xt[2].reshape(-1)]))
        s["ytab"] = table(("x", ih, nh), lambda: np.concatenate([yt[0], yt[1], yt[2].reshape(-1)]))
        s["xnn"] = table(("n", iw, nw), lambda: nearest_table(iw, nw))
        s["ynn"] = table(("n", ih, nh), lambda: nearest_table(ih, nh))
        if p["rotate"]:
            rot = int(p["rotation"])
            s["rot"] = table(("r", rot), lambda: rotation_tables(w, h, rot))
        if iw != nw:
            s["tmp_off"] = tmp_off
            tmp_off += (ih * nw * 3 + 15) // 16 * 16
            max_elems = max(max_elems, ih * nw)
        if p.get("r") is None:
            s["lut"] = -1
        else:
            s["lut"] = i * 768
            luts[i] = hsv_luts(p["r"])
    tables = np.concatenate(tab_parts) if tab_parts else np.zeros(1, dtype=np.int32)
    if tables.size >= 2 ** 31 or src_off >= 2 ** 40:
        raise ValueError("augmentation batch too large")
    return AugmentPlan(torch.from_numpy(samples.view(np.uint8).reshape(n, AUG_SAMPLE.itemsize).copy()),
                       torch.from_numpy(np.concatenate(src_parts)), torch.from_numpy(tables), torch.from_numpy(luts.reshape(-1)),
                       (h, w), int(max_elems), int(tmp_off))


# ------------------------------------------------------------------------------------------------ device side
class DeviceAugmenter:
    """Runs packed batches through the four augmentation kernels on one device."""

    def __init__(self, device=None):
        from ..backend import get_backend
        self.B = get_backend()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.cubic = torch.from_numpy(cubic_weights()).to(self.device)

    def upload(self, plan: AugmentPlan, non_blocking: bool = True):
        """Host blobs -> device (four copies; pinned blobs make them asynchronous)."""
        return tuple(t.to(self.device, non_blocking=non_blocking) for t in (plan.samples, plan.src, plan.tables, plan.luts))

    def run(self, plan: AugmentPlan, blobs=None, out=None):
        """-> (uint8 [B,H,W,3] pixels, uint8 [B,H,W] class maps) on the device."""
        samples, src, tables, luts = blobs if blobs is not None else self.upload(plan)
        h, w = plan.shape
        return self.B.augment_batch(samples, src, tables, luts, self.cubic, len(plan), h, w, plan.max_elems, plan.tmp_bytes,
                                    (w // CV_VECTOR_PIXELS) * CV_VECTOR_PIXELS, out)


class DeviceAugmentLoader:
    """Wraps an iterable of ``AugmentPlan`` batches (a DataLoader with ``deeplab_dataset_collate``): uploads batch i+1 on
    a copy stream while batch i trains, runs the kernels on the consumer's stream and yields
    ``(pixels uint8 [B,H,W,3], class maps uint8 [B,H,W], None)`` - the batch triple of fit_one_epoch with the one-hot
    labels left implicit.  Two sets of device blobs are used alternately; a set is overwritten only after the kernels
    that read it have run."""

    on_device = True

    def __init__(self, plans, device=None):
        self.plans = plans
        self.aug = DeviceAugmenter(device)
        self.copy_stream = torch.cuda.Stream(self.aug.device)

    def __len__(self):
        return len(self.plans)

    def __iter__(self):
        dev = self.aug.device
        it = iter(self.plans)
        done = [None, None]                  # event: kernels of the batch that used blob set s have been enqueued and run
        k = 0

        def fetch():
            nonlocal k
            try:
                plan = next(it)
            except StopIteration:
                return None
            s = k & 1
            k += 1
            with torch.cuda.stream(self.copy_stream):
                if done[s] is not None:
                    self.copy_stream.wait_event(done[s])
                blobs = self.aug.upload(plan)
                ready = torch.cuda.Event()
                ready.record(self.copy_stream)
            return plan, blobs, ready, s

        nxt = fetch()
        while nxt is not None:
            plan, blobs, ready, s = nxt
            nxt = fetch()
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ready)
            imgs, labs = self.aug.run(plan, blobs)
            for b in blobs:
                b.record_stream(cur)
            ev = torch.cuda.Event()
            ev.record(cur)
            done[s] = ev
            yield imgs, labs, None


# ------------------------------------------------------------------------------------------------ the reference's names
def _cvt_rgb(image):
    """utils.utils.cvtColor: anything that is not already an RGB image is converted."""
    if len(np.shape(image)) == 3 and np.shape(image)[2] == 3:
        return image
    return image.convert("RGB")


class DeeplabDataset(torch.utils.data.Dataset):
    """Same constructor and file layout as the reference (dataloader.py:12-35).  ``__getitem__`` returns the decoded
    sources and the drawn decisions instead of an augmented float image: the arithmetic runs on the device."""

    def __init__(self, annotation_lines, input_shape, num_classes, train, dataset_path):
        super().__init__()
        self.annotation_lines = annotation_lines
        self.length = len(annotation_lines)
        self.input_shape = input_shape
        self.num_classes = num_classes
        self.train = train
        self.dataset_path = dataset_path

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        from PIL import Image
        name = self.annotation_lines[index].split()[0]
        jpg = Image.open(os.path.join(os.path.join(self.dataset_path, "VOC2007/JPEGImages"), name + ".jpg"))
        png = Image.open(os.path.join(os.path.join(self.dataset_path, "VOC2007/SegmentationClass"), name + ".png"))
        return self.decode(jpg, png)

    def decode(self, image, label):
        """PIL image + PIL class map -> (uint8 [ih,iw,3], uint8 [ih,iw], decisions, canvas shape)."""
        img = np.asarray(_cvt_rgb(image), dtype=np.uint8)
        lab = np.asarray(label, dtype=np.uint8)
        ih, iw = lab.shape[:2]
        p = draw_params(iw, ih, self.input_shape) if self.train else letterbox_params(iw, ih, self.input_shape)
        return img, lab, p, tuple(int(v) for v in self.input_shape)

    def rand(self, a=0, b=1):
        return np.random.rand() * (b - a) + a

    def get_random_data(self, image, label, input_shape, jitter=.3, hue=.1, sat=0.7, val=0.3, random=True, device=None):
        """The reference's method (dataloader.py:55-154) for ONE image, computed on the device: returns the augmented
        uint8 RGB array ``[h, w, 3]`` and the uint8 class map ``[h, w]``.  Consumes ``np.random`` exactly as the reference."""
        img = np.asarray(_cvt_rgb(image), dtype=np.uint8)
        lab = np.asarray(label, dtype=np.uint8)
        ih, iw = lab.shape[:2]
        p = draw_params(iw, ih, input_shape, np.random, jitter, hue, sat, val) if random else letterbox_params(iw, ih, input_shape)
        aug = getattr(self, "_augmenter", None)
        if aug is None or (device is not None and torch.device(device) != aug.device):
            aug = self._augmenter = DeviceAugmenter(device)
        imgs, labs = aug.run(pack_batch([(img, lab, p)], input_shape))
        return imgs[0].cpu().numpy(), labs[0].cpu().numpy()


def deeplab_dataset_collate(batch) -> AugmentPlan:
    """DataLoader ``collate_fn`` (reference: dataloader.py:158-169): packs the decoded samples of a batch."""
    shape = batch[0][3]
    return pack_batch([(img, lab, p) for img, lab, p, _ in batch], shape)
