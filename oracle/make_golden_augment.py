"""Golden vectors of the reference's training augmentation (TEST INFRASTRUCTURE; run in the build container only).

    python oracle/make_golden_augment.py        ->  tests/golden/augment.npz

Runs the UNMODIFIED ``DeeplabDataset.get_random_data`` of /root/reference (Segmentation/deeplabv3+/utils/dataloader.py
:55-154) under fixed numpy seeds on small synthetic images and stores inputs, seeds and outputs.  The seeds are chosen
by scanning so that the set covers: both aspect branches, enlarging beyond the canvas (negative paste offsets) and
reducing, flip, blur, rotation, every combination of blur x rotation, the deterministic validation path (random=False),
and a canvas whose width is not a multiple of OpenCV's 32-pixel vectors (scalar tail of the HSV conversion)."""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/Segmentation/deeplabv3+")

from utils.dataloader import DeeplabDataset  # noqa: E402  (the reference's)

from oracle import augment_ref as A  # noqa: E402


def synth(rng, ih, iw):
    """A smooth-ish RGB image with texture plus a blocky class map (values 0..5 and a few 255 'white border' pixels)."""
    low = rng.randint(0, 256, (ih // 6 + 2, iw // 6 + 2, 3)).astype(np.uint8)
    img = np.asarray(Image.fromarray(low).resize((iw, ih), Image.BILINEAR)).astype(np.int64)
    img = np.clip(img + rng.randint(-25, 26, img.shape), 0, 255).astype(np.uint8)
    lab = rng.randint(0, 6, (ih // 9 + 1, iw // 9 + 1)).astype(np.uint8)
    lab = np.repeat(np.repeat(lab, 9, 0), 9, 1)[:ih, :iw].copy()
    lab[rng.rand(ih, iw) < 0.01] = 255
    return img, lab


def main():
    ds = DeeplabDataset(["x"], (96, 96), 5, True, "/nonexistent")
    rng = np.random.RandomState(1234)
    cases = []
    want = [dict(blur=False, rotate=False), dict(blur=True, rotate=False), dict(blur=False, rotate=True),
            dict(blur=True, rotate=True), dict(flip=True, big=True), dict(flip=False, big=False, tall=True),
            dict(big=True, rotate=True), dict(tall=False, blur=True)]
    shapes = [((96, 96), (57, 83)), ((96, 96), (120, 70)), ((64, 128), (90, 90)), ((96, 96), (75, 101)),
              ((96, 96), (66, 49)), ((128, 96), (140, 100)), ((96, 96), (80, 60)), ((80, 72), (77, 91))]
    seed = 0
    for wanted, (shape, (ih, iw)) in zip(want, shapes):
        while True:
            seed += 1
            np.random.seed(seed)
            p = A.draw_params(iw, ih, shape, np.random)
            props = dict(blur=p["blur"], rotate=p["rotate"], flip=p["flip"], big=p["nw"] > shape[1] or p["nh"] > shape[0],
                         tall=p["nh"] > p["nw"])
            if all(props[k] == v for k, v in wanted.items()) and p["nw"] > 0 and p["nh"] > 0:
                break
        img, lab = synth(rng, ih, iw)
        np.random.seed(seed)
        out_img, out_lab = ds.get_random_data(Image.fromarray(img), Image.fromarray(lab), shape, random=True)
        cases.append(dict(seed=seed, shape=shape, img=img, lab=lab, out_img=np.asarray(out_img, np.uint8),
                          out_lab=np.asarray(out_lab, np.uint8), random=True))
        print("case", len(cases) - 1, "seed", seed, "canvas", shape, "source", (ih, iw), {k: (v if not isinstance(v, np.ndarray) else v.round(3).tolist()) for k, v in p.items()})
    for shape, (ih, iw) in (((96, 96), (50, 83)), ((64, 96), (130, 70))):
        img, lab = synth(rng, ih, iw)
        out_img, out_lab = ds.get_random_data(Image.fromarray(img), Image.fromarray(lab), shape, random=False)
        cases.append(dict(seed=-1, shape=shape, img=img, lab=lab, out_img=np.asarray(out_img, np.uint8),
                          out_lab=np.asarray(out_lab, np.uint8), random=False))
    # the restatement must reproduce every case before the vectors are written
    for i, c in enumerate(cases):
        ih, iw = c["lab"].shape
        if c["random"]:
            np.random.seed(c["seed"])
            p = A.draw_params(iw, ih, c["shape"], np.random)
        else:
            p = A.letterbox_params(iw, ih, c["shape"])
        got_img, got_lab = A.apply_params(c["img"], c["lab"], c["shape"], p)
        assert np.array_equal(got_img, c["out_img"]), ("image", i, int((got_img != c["out_img"]).sum()))
        assert np.array_equal(got_lab, c["out_lab"]), ("label", i)
    out = {"n": np.int64(len(cases))}
    for i, c in enumerate(cases):
        out["seed_%d" % i] = np.int64(c["seed"])
        out["shape_%d" % i] = np.asarray(c["shape"], np.int64)
        out["random_%d" % i] = np.int64(c["random"])
        for k in ("img", "lab", "out_img", "out_lab"):
            out["%s_%d" % (k, i)] = c[k]
    path = os.path.join(ROOT, "tests", "golden", "augment.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(cases), "cases: the restatement reproduces all of them bit for bit")


if __name__ == "__main__":
    main()
