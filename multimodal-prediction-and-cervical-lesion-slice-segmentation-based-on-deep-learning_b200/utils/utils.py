"""Small host helpers of the reference's ``utils/utils.py`` that its train / predict entry points import
(cvtColor :11, resize_image :21 letterbox, get_lr :39, seed_everything :46, preprocess_input :64, show_config :68).
Pure host code (PIL / numpy); the letterbox geometry is what ``deeplab.DeeplabV3`` feeds the crop of the fused
post-processing kernel with."""
import random

import numpy as np
import torch
from PIL import Image


def cvtColor(image):
    """PIL image -> RGB (grey / RGBA inputs are converted, RGB passes through)."""
    shape = np.shape(image)
    if len(shape) == 3 and shape[2] == 3:
        return image
    return image.convert("RGB")


def letterbox_geometry(iw, ih, w, h):
    """(nw, nh, left, top): the aspect-preserving fit of an iw x ih image into w x h and where it is pasted."""
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    return nw, nh, (w - nw) // 2, (h - nh) // 2


def resize_image(image, size):
    """Letterbox ``image`` into ``size`` = (w, h) on a grey (128) canvas; returns (canvas, nw, nh)."""
    w, h = size
    nw, nh, left, top = letterbox_geometry(image.size[0], image.size[1], w, h)
    canvas = Image.new("RGB", (w, h), (128, 128, 128))
    canvas.paste(image.resize((nw, nh), Image.BICUBIC), (left, top))
    return canvas, nw, nh


def get_lr(optimizer):
    for group in optimizer.param_groups:
        return group["lr"]


def seed_everything(seed=11):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def worker_init_fn(worker_id, rank, seed):
    worker_seed = rank + seed
    random.seed(worker_seed)
    np.random.seed(worker_seed)
    torch.manual_seed(worker_seed)


def preprocess_input(image):
    """uint8-range float image -> [0, 1] (the only normalisation of the segmentation input contract)."""
    image /= 255.0
    return image


def show_config(**kwargs):
    line = "-" * 70
    print("Configurations:")
    print(line)
    print("|%25s | %40s|" % ("keys", "values"))
    print(line)
    for key, value in kwargs.items():
        print("|%25s | %40s|" % (str(key), str(value)))
    print(line)
