"""CPU tests: the oracle restatement (oracle/) against the golden vectors that
oracle/make_golden.py produced from the UNMODIFIED reference modules."""
import os

import numpy as np
import pytest
import torch

from oracle import deeplab_ref as O
from oracle import losses_ref as L
from oracle.make_golden import GRAD_KEYS, STAT_KEYS, subsample

CLS_W = torch.tensor([1, 1, 5, 3, 4], dtype=torch.float32)


def relerr(a, b):
    a = torch.as_tensor(a); b = torch.as_tensor(b)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("backbone,ds", [("xception", 16), ("xception", 8), ("mobilenet", 16), ("mobilenet", 8)])
def test_eval_logits_match_reference(golden_dir, backbone, ds):
    g = np.load(os.path.join(golden_dir, f"eval_{backbone}_ds{ds}.npz"))
    state = O.make_state(backbone, 5, ds, seed=int(g["seed"]))
    # the synthetic-weight generator must reproduce what the fixture was made with
    assert np.array_equal(state["cls_conv.weight"].numpy(), g["w_probe"])
    with torch.no_grad():
        y = O.deeplab_forward(torch.from_numpy(g["imgs"]), state, backbone, ds, False)
    assert relerr(y, g["logits"]) < 1e-5
    assert (y.argmax(1).numpy() == g["logits"].argmax(1)).mean() >= 0.999


@pytest.mark.parametrize("backbone", ["xception", "mobilenet"])
def test_train_step_matches_reference(golden_dir, backbone):
    g = np.load(os.path.join(golden_dir, f"train_{backbone}.npz"))
    state = O.make_state(backbone, 5, 16, seed=int(g["seed"]))
    st = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
          for k, v in state.items()}
    imgs = torch.from_numpy(g["imgs"]); pngs = torch.from_numpy(g["pngs"])
    labels = torch.eye(6)[pngs]
    y = O.deeplab_forward(imgs, st, backbone, 16, True, dropout=False)
    assert relerr(y.detach(), g["logits"]) < 1e-4
    focal = L.focal_loss(y, pngs, CLS_W, 5); dice = L.dice_loss(y, labels)
    assert abs(float(focal) - float(g["focal"])) < 1e-4
    assert abs(float(dice) - float(g["dice"])) < 1e-5
    assert abs(float(L.ce_loss(y, pngs, CLS_W, 5)) - float(g["ce"])) < 1e-4
    assert abs(float(L.f_score(y.detach(), labels)) - float(g["f_score"])) < 1e-5
    (focal + dice).backward()
    for k in GRAD_KEYS[backbone]:
        assert relerr(subsample(st[k].grad), g["grad:" + k]) < 2e-3, k
    for k in STAT_KEYS[backbone]:
        assert relerr(st[k], g["stat:" + k]) < 1e-5, k


def test_losses_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    pngs = torch.from_numpy(g["pngs"]); labels = torch.eye(6)[pngs]
    for name, fn in (("ce", lambda z: L.ce_loss(z, pngs, CLS_W, 5)),
                     ("focal", lambda z: L.focal_loss(z, pngs, CLS_W, 5)),
                     ("dice", lambda z: L.dice_loss(z, labels))):
        z = torch.from_numpy(g["logits"]).clone().requires_grad_(True)
        v = fn(z); v.backward()
        assert abs(float(v) - float(g[name])) < 1e-6
        assert np.abs(z.grad.numpy() - g["d" + name]).max() < 1e-7
    z = torch.from_numpy(g["logits"])
    assert abs(float(L.f_score(z, labels)) - float(g["f_score"])) < 1e-6
    zl = z[:, :, :6, :10].clone().requires_grad_(True)
    v = L.focal_loss(zl, pngs, CLS_W, 5) + L.dice_loss(zl, labels); v.backward()
    assert abs(float(v) - float(g["focal_dice_lowres"])) < 1e-5
    assert np.abs(zl.grad.numpy() - g["dlowres"]).max() < 1e-6


def test_lr_schedule_matches_reference(golden_dir):
    rows = np.load(os.path.join(golden_dir, "lr_schedule.npz"))["rows"]
    for kind, it, val in rows:
        assert abs(L.lr_at("cos" if kind == 0 else "step", 5e-4, 5e-6, 100, int(it)) - val) < 1e-12


def test_state_schema_counts():
    assert len(O.state_schema("xception")) == 857
    assert len(O.state_schema("mobilenet")) == 371
    n = sum(int(np.prod(s)) for k, s, kind in O.state_schema("xception") if kind in ("conv", "cbias", "gamma", "beta"))
    assert n == 54_709_445
    n = sum(int(np.prod(s)) for k, s, kind in O.state_schema("mobilenet") if kind in ("conv", "cbias", "gamma", "beta"))
    assert n == 5_814_037
