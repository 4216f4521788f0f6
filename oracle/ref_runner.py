"""Drive the UNMODIFIED reference (installed under baseline/_ref by oracle/build_ref.py) on synthetic batches.

TEST / BENCH INFRASTRUCTURE ONLY - the product never imports this.

``time_fit_one_epoch`` calls the reference's own training entry point for the segmentation path,
``utils.utils_fit.fit_one_epoch`` (Segmentation/deeplabv3+/utils/utils_fit.py:31-198), with the reference's own model
(``nets.deeplabv3_plus.DeepLab``, Xception ds=16, ``weights_init`` as train.py:314), optimizer (Adam, train.py:472-476),
``GradScaler`` + fp16 autocast when ``fp16`` (train.py:82,  utils_fit.py:92-121), objective flags of the script
(``dice_loss=True, focal_loss=True``, train.py:259-265) and class weights (train.py:274).  The data generator yields
synthetic host batches in the loader's contract (dataloader.py:158-169: fp32 NCHW images in [0,1], int64 class maps,
fp32 one-hot labels) and records a timestamp at every hand-over, so the per-step time contains everything the reference
does per step: the host->device copies (``imgs.cuda(local_rank)``), forward, loss, backward, optimizer and the two
``.item()`` syncs.

CLI (used by bench.py in a subprocess so the reference's top-level ``nets`` / ``utils`` packages never meet the
product's):   python oracle/ref_runner.py --device cuda --batch 32 --steps 5 --warmup 2 --fp16 [--channels-last]
prints one JSON line."""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SEG = os.path.join(os.path.dirname(HERE), "baseline", "_ref", "seg")


def import_reference_seg():
    """Put baseline/_ref/seg first on sys.path and import the reference modules as the scripts do."""
    if not os.path.exists(os.path.join(SEG, "utils", "utils_fit.py")):
        raise RuntimeError("baseline/_ref is not installed (run oracle/build_ref.py where /root/reference exists)")
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.pyplot"):     # imported for plots that are never drawn here
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod = types.ModuleType(name)
                mod.use = lambda *a, **k: None
                sys.modules[name] = mod
    if "matplotlib.pyplot" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for k in [k for k in sys.modules if k == "nets" or k.startswith("nets.") or k == "utils" or k.startswith("utils.")]:
        del sys.modules[k]
    sys.path.insert(0, SEG)
    import nets.deeplabv3_plus as dl
    import nets.deeplabv3_training as tr
    import utils.utils_fit as fit
    return dl, tr, fit


class _History:
    val_loss: list = []

    def append_loss(self, *a):
        pass


class _Eval:
    def on_epoch_end(self, *a):
        pass


def time_fit_one_epoch(device: str, batch: int, size: int, steps: int, warmup: int, fp16: bool, channels_last: bool = False,
                       threads: int | None = None, budget_s: float | None = None):
    import numpy as np
    import torch
    dl, tr, fit = import_reference_seg()
    cuda = device == "cuda"
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = dl.DeepLab(num_classes=5, backbone="xception", downsample_factor=16, pretrained=False)
    with contextlib.redirect_stdout(io.StringIO()):
        tr.weights_init(model)
    if cuda:
        model = model.cuda()
        torch.backends.cudnn.benchmark = True       # train.py:383
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    model_train = model.train()
    opt = torch.optim.Adam(model.parameters(), 1e-4, betas=(0.9, 0.999), weight_decay=0)
    scaler = torch.cuda.amp.GradScaler() if fp16 else None
    g = torch.Generator().manual_seed(1000)
    imgs = torch.rand(batch, 3, size, size, generator=g)
    pngs = torch.randint(0, 5, (batch, size, size), generator=g)
    pngs = torch.where(torch.rand(batch, size, size, generator=g) < 0.01, torch.full_like(pngs, 5), pngs)
    labels = torch.eye(6)[pngs]
    if cuda:
        imgs, pngs, labels = imgs.pin_memory(), pngs.pin_memory(), labels.pin_memory()
    if channels_last:
        imgs = imgs.contiguous(memory_format=torch.channels_last)
    stamps = []
    t_begin = time.perf_counter()

    def gen():
        for i in range(warmup + steps):
            if cuda:
                torch.cuda.synchronize()
            stamps.append(time.perf_counter())
            if budget_s is not None and i > warmup and stamps[-1] - t_begin > budget_s:
                return
            yield imgs, pngs, labels
        if cuda:
            torch.cuda.synchronize()
        stamps.append(time.perf_counter())

    cls_weights = np.array([1, 1, 5, 3, 4], np.float32)
    total = warmup + steps
    with tempfile.TemporaryDirectory() as save_dir, contextlib.redirect_stdout(io.StringIO()), \
            contextlib.redirect_stderr(io.StringIO()):
        fit.fit_one_epoch(model_train, model, _History(), _Eval(), opt, 0, total, 1, gen(), [(imgs[:1], pngs[:1], labels[:1])],
                          2, cuda, True, True, cls_weights, 5, fp16, scaler, 1000, save_dir, 0)
    d = [b - a for a, b in zip(stamps[:-1], stamps[1:])][warmup:]
    sec = sum(d) / len(d)
    return {"images_per_s": batch / sec, "ms_per_step": sec * 1e3, "steps_timed": len(d), "batch": batch, "size": size,
            "fp16_autocast": bool(fp16), "channels_last": bool(channels_last), "device": device,
            "torch_threads": torch.get_num_threads()}


def augment_sources(count: int, seed: int = 0):
    """VOC-sized synthetic decoded images (375 x 500 landscape / 500 x 375 portrait) with class maps - the same
    sources bench.py's augmentation leg packs for the device."""
    import numpy as np
    from PIL import Image
    rng = np.random.RandomState(seed)
    out = []
    for i in range(count):
        ih, iw = (375, 500) if i % 2 == 0 else (500, 375)
        low = rng.randint(0, 256, (ih // 6 + 2, iw // 6 + 2, 3)).astype(np.uint8)
        img = np.asarray(Image.fromarray(low).resize((iw, ih), Image.BILINEAR))
        out.append((img, rng.randint(0, 6, (ih, iw)).astype(np.uint8)))
    return out


def time_augment(count: int, size: int):
    """The unmodified reference's DeeplabDataset.get_random_data + loader tail (dataloader.py:36-49) on one host core,
    images per second (one DataLoader worker; train.py:281 starts four)."""
    import time

    import numpy as np
    from PIL import Image
    import_reference_seg()
    from utils.dataloader import DeeplabDataset
    from utils.utils import preprocess_input
    ds = DeeplabDataset(["x"], (size, size), 5, True, "/nonexistent")
    srcs = [(Image.fromarray(a), Image.fromarray(b)) for a, b in augment_sources(count)]
    np.random.seed(0)
    t0 = time.perf_counter()
    for jpg, png in srcs:
        jpg, png = ds.get_random_data(jpg, png, (size, size), random=True)
        jpg = np.transpose(preprocess_input(np.array(jpg, np.float64)), [2, 0, 1])
        png = np.array(png)
        png[png >= 5] = 5
        np.eye(6)[png.reshape([-1])].reshape((size, size, 6))
    sec = time.perf_counter() - t0
    return {"images_per_s": count / sec, "ms_per_image": sec / count * 1e3, "images": count}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--augment", type=int, default=0, help="time get_random_data on this many images instead of training")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--fp16", action="store_true")
    ap.add_argument("--channels-last", action="store_true")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--budget-s", type=float, default=0.0)
    a = ap.parse_args()
    if a.augment:
        print(json.dumps(time_augment(a.augment, a.size)), flush=True)
        return
    out = time_fit_one_epoch(a.device, a.batch, a.size, a.steps, a.warmup, a.fp16, a.channels_last, a.threads or None,
                             a.budget_s or None)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
