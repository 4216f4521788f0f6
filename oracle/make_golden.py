"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

For every case it (1) instantiates the reference ``DeepLab`` (nets/deeplabv3_plus.py:116)
and loads the deterministic synthetic weights of ``oracle.deeplab_ref.make_state`` with
``strict=True`` (which also pins the 857/371-entry state_dict schema), (2) runs the
reference forward / losses / backward, (3) runs the oracle restatement on the same inputs
and asserts agreement, and (4) stores inputs + reference outputs as small fixtures.  The
fixtures travel to the GPU box; /root/reference does not.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/Segmentation/deeplabv3+"

from oracle import deeplab_ref as O  # noqa: E402
from oracle import losses_ref as L  # noqa: E402


def _import_reference():
    # matplotlib is absent; utils_metrics only needs the name to import
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib"); mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
    sys.path.insert(0, REF)
    from nets.deeplabv3_plus import DeepLab
    from nets.deeplabv3_training import CE_Loss, Focal_Loss, Dice_loss
    from utils.utils_metrics import f_score
    sys.path.remove(REF)
    # drop the reference's top-level package names so they never shadow the product's
    for k in [k for k in sys.modules if k == "nets" or k.startswith("nets.") or k == "utils" or k.startswith("utils.")]:
        sys.modules.pop(k)
    return DeepLab, CE_Loss, Focal_Loss, Dice_loss, f_score


GRAD_KEYS = {
    "xception": ["cls_conv.weight", "cls_conv.bias", "cat_conv.0.weight", "aspp.branch3.0.weight",
                 "aspp.branch5_conv.weight", "backbone.block20.skip.weight",
                 "backbone.block7.sepconv2.pointwise.weight", "backbone.block7.sepconv2.depthwise.weight",
                 "backbone.block2.sepconv2.bn2.weight", "backbone.block1.skipbn.bias",
                 "backbone.conv2.weight", "backbone.conv1.weight", "shortcut_conv.0.weight"],
    "mobilenet": ["cls_conv.weight", "cat_conv.4.weight", "aspp.branch2.0.weight",
                  "backbone.features.17.conv.6.weight", "backbone.features.15.conv.3.weight",
                  "backbone.features.8.conv.0.weight", "backbone.features.3.conv.7.weight",
                  "backbone.features.1.conv.0.weight", "backbone.features.0.0.weight",
                  "backbone.features.0.1.weight", "shortcut_conv.0.weight"],
}
STAT_KEYS = {
    "xception": ["backbone.bn1.running_mean", "backbone.bn1.running_var",
                 "backbone.block12.sepconv1.bn2.running_var", "aspp.branch5_bn.running_mean",
                 "cat_conv.5.running_var"],
    "mobilenet": ["backbone.features.0.1.running_mean", "backbone.features.9.conv.4.running_var",
                  "aspp.conv_cat.1.running_mean", "cat_conv.1.running_var"],
}


def is_null_grad_param(key: str) -> bool:
    """Parameters whose training-mode gradient is exactly zero in exact arithmetic: a
    per-channel constant added right before a batch-statistics BatchNorm (possibly through
    a linear 1x1 conv).  Xception sepconv ``bn1.bias`` with activate_first (blocks 1-20),
    and every biased conv of ASPP / decoder that feeds a BN."""
    import re
    if re.match(r"backbone\.block\d+\.sepconv\d\.bn1\.bias$", key):
        return True
    # MobileNetV2: the closing (linear) BN bias of every inverted-residual block only ever
    # reaches 1x1 convs that are followed by batch-stat BNs (through the residual adds too)
    if re.match(r"backbone\.features\.\d+\.conv\.7\.bias$", key) or key == "backbone.features.1.conv.4.bias":
        return True
    return key in ("aspp.branch1.0.bias", "aspp.branch2.0.bias", "aspp.branch3.0.bias",
                   "aspp.branch4.0.bias", "aspp.branch5_conv.bias", "aspp.conv_cat.0.bias",
                   "shortcut_conv.0.bias", "cat_conv.0.bias", "cat_conv.4.bias")


def subsample(t, limit: int = 16384):
    """Fixtures keep at most ~``limit`` evenly strided samples of a large tensor."""
    flat = t.reshape(-1)
    stride = flat.numel() // limit + 1
    return flat[::stride]


def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    DeepLab, CE_Loss, Focal_Loss, Dice_loss, f_score = _import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    cls_w = torch.tensor([1, 1, 5, 3, 4], dtype=torch.float32)  # train.py:274

    # ------------------------------------------------------------------ eval-mode logits
    for backbone in ("xception", "mobilenet"):
        for ds in (16, 8):
            size = 96 if ds == 16 else 64
            state = O.make_state(backbone, 5, ds, seed=7)
            ref = DeepLab(5, backbone, False, ds)
            ref.load_state_dict(state, strict=True)
            ref.eval()
            imgs, _, _ = O.synthetic_batch(1, size, seed=3)
            with torch.no_grad():
                y_ref = ref(imgs)
                y_or, low_or = O.deeplab_forward(imgs, state, backbone, ds, False, return_lowres=True)
            e = relerr(y_or, y_ref)
            print(f"eval {backbone} ds={ds} {size}px: oracle-vs-reference rel err {e:.2e}")
            assert e < 1e-5, e
            np.savez_compressed(os.path.join(out_dir, f"eval_{backbone}_ds{ds}.npz"),
                                seed=7, size=size, imgs=imgs.numpy(), logits=y_ref.numpy(),
                                w_probe=state["cls_conv.weight"].numpy())

    # ------------------------------------------------------------------ train step (fwd+loss+bwd)
    for backbone in ("xception", "mobilenet"):
        ds, size, bsz = 16, 64, 2
        state = O.make_state(backbone, 5, ds, seed=11)
        ref = DeepLab(5, backbone, False, ds)
        ref.load_state_dict(state, strict=True)
        ref.train()
        for m in ref.modules():  # dropout off for exact parity
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        imgs, pngs, labels = O.synthetic_batch(bsz, size, seed=5)
        y_ref = ref(imgs)
        focal = Focal_Loss(y_ref, pngs, cls_w, num_classes=5)
        dice = Dice_loss(y_ref, labels)
        ce = CE_Loss(y_ref, pngs, cls_w, num_classes=5)
        with torch.no_grad():
            fs = f_score(y_ref, labels)
        (focal + dice).backward()
        ref_grads = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
        ref_sd = ref.state_dict()

        st = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
              for k, v in state.items()}
        y_or = O.deeplab_forward(imgs, st, backbone, ds, True, dropout=False)
        focal_o = L.focal_loss(y_or, pngs, cls_w, 5)
        dice_o = L.dice_loss(y_or, labels)
        ce_o = L.ce_loss(y_or, pngs, cls_w, 5)
        fs_o = L.f_score(y_or.detach(), labels)
        (focal_o + dice_o).backward()
        print(f"train {backbone}: logits rel err {relerr(y_or.detach(), y_ref.detach()):.2e}; "
              f"focal {float(focal):.6f}/{float(focal_o):.6f} dice {float(dice):.6f}/{float(dice_o):.6f} "
              f"ce {float(ce):.6f}/{float(ce_o):.6f} f {float(fs):.6f}/{float(fs_o):.6f}")
        assert relerr(y_or.detach(), y_ref.detach()) < 1e-4
        for a, b in ((focal, focal_o), (dice, dice_o), (ce, ce_o), (fs, fs_o)):
            assert abs(float(a) - float(b)) <= 1e-5 * max(1.0, abs(float(a)))
        worst, worst_null = 0.0, 0.0
        for k, g in ref_grads.items():
            if is_null_grad_param(k):
                # mathematically zero gradient (a constant shift in front of a batch-stat BN):
                # both sides hold rounding noise only
                worst_null = max(worst_null, float(g.abs().max()), float(st[k].grad.abs().max()))
            else:
                worst = max(worst, relerr(st[k].grad, g))
        print(f"  worst grad rel err over {len(ref_grads)} params: {worst:.2e} (null-grad noise {worst_null:.1e})")
        assert worst < 2e-3, worst
        assert worst_null < 1e-3, worst_null
        for k in STAT_KEYS[backbone]:
            assert relerr(st[k], ref_sd[k]) < 1e-5, k
        payload = dict(seed=11, size=size, imgs=imgs.numpy(), pngs=pngs.numpy().astype(np.int64),
                       logits=y_ref.detach().numpy(), focal=float(focal), dice=float(dice), ce=float(ce),
                       f_score=float(fs))
        for k in GRAD_KEYS[backbone]:
            payload["grad:" + k] = subsample(ref_grads[k]).numpy()
        for k in STAT_KEYS[backbone]:
            payload["stat:" + k] = ref_sd[k].numpy()
        np.savez_compressed(os.path.join(out_dir, f"train_{backbone}.npz"), **payload)

    # ------------------------------------------------------------------ losses on raw logits
    g = torch.Generator().manual_seed(42)
    logits = (3.0 * torch.randn(2, 5, 24, 40, generator=g)).requires_grad_(True)
    _, pngs, labels = O.synthetic_batch(2, 40, seed=9, ignore_frac=0.05)
    pngs, labels = pngs[:, :24].contiguous(), labels[:, :24].contiguous()
    res = {}
    for name, fn in (("ce", lambda z: CE_Loss(z, pngs, cls_w, 5)),
                     ("focal", lambda z: Focal_Loss(z, pngs, cls_w, 5)),
                     ("dice", lambda z: Dice_loss(z, labels))):
        z = logits.detach().clone().requires_grad_(True)
        v = fn(z); v.backward()
        res[name] = float(v); res["d" + name] = z.grad.numpy()
    with torch.no_grad():
        res["f_score"] = float(f_score(logits, labels))
    # low-res logits + target at 4x size (exercises the in-loss bilinear resize)
    zl = logits.detach()[:, :, :6, :10].clone().requires_grad_(True)
    v = Focal_Loss(zl, pngs, cls_w, 5) + Dice_loss(zl, labels); v.backward()
    res["focal_dice_lowres"] = float(v); res["dlowres"] = zl.grad.numpy()
    for name, fn in (("ce", lambda z: L.ce_loss(z, pngs, cls_w, 5)),
                     ("focal", lambda z: L.focal_loss(z, pngs, cls_w, 5)),
                     ("dice", lambda z: L.dice_loss(z, labels))):
        z = logits.detach().clone().requires_grad_(True)
        v = fn(z); v.backward()
        assert abs(float(v) - res[name]) < 1e-6, name
        assert np.abs(z.grad.numpy() - res["d" + name]).max() < 1e-7, name
    assert abs(float(L.f_score(logits.detach(), labels)) - res["f_score"]) < 1e-6
    np.savez_compressed(os.path.join(out_dir, "losses.npz"), logits=logits.detach().numpy(),
                        pngs=pngs.numpy(), **res)
    print("losses ok:", {k: v for k, v in res.items() if isinstance(v, float)})

    # ------------------------------------------------------------------ lr schedule
    sys.path.insert(0, REF)
    from nets.deeplabv3_training import get_lr_scheduler
    sys.path.remove(REF)
    rows = []
    for kind in ("cos", "step"):
        f = get_lr_scheduler(kind, 5e-4, 5e-6, 100)
        for it in (0, 1, 2, 3, 4, 10, 50, 84, 85, 99):
            rows.append((0 if kind == "cos" else 1, it, f(it)))
            assert abs(L.lr_at(kind, 5e-4, 5e-6, 100, it) - f(it)) < 1e-12
    np.savez_compressed(os.path.join(out_dir, "lr_schedule.npz"), rows=np.array(rows, dtype=np.float64))
    print("golden vectors written to", out_dir)


if __name__ == "__main__":
    main()
