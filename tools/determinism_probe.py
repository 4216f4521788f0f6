"""Run-to-run spread of ONE train-mode forward+backward from identical state (GPU box only).

    python tools/determinism_probe.py [--reps 6]

For each (dtype, image size, batch) the same model state and batch are evaluated `reps` times; prints the focal / dice
loss of every repeat, the relative spread (max - min) / mean, and the relative spread of the gradient norm.  Used to
answer VERDICT r01 "what's weak" #2: is the bf16 train-mode spread seen in smoke() (64x64, batch 4) a rounding cascade
through batch-statistics BatchNorm over tiny populations (then it must shrink with the population) or a race (then it
would not)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from cervix_b200.nets.deeplabv3_plus import DeepLab
from cervix_b200.nets.deeplabv3_training import seg_objective
from oracle import deeplab_ref as O

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=6)
ap.add_argument("--cases", default="f32:64:4,bf16:64:4,bf16:128:4,bf16:256:8,bf16:512:8,bf16:512:32")
args = ap.parse_args()

cls_w = torch.tensor([1, 1, 5, 3, 4], dtype=torch.float32).cuda()
state = O.make_state("xception", 5, 16, seed=3)
for case in args.cases.split(","):
    dt, size, bsz = case.split(":")
    size, bsz = int(size), int(bsz)
    dtype = torch.float32 if dt == "f32" else torch.bfloat16
    imgs, pngs, labels = [t.cuda() for t in O.synthetic_batch(bsz, size, seed=1)]
    model = DeepLab(5, "xception", False, 16).set_compute_dtype(dtype)
    model.load_state_dict(state)
    model.cuda().train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    focals, dices, gnorms = [], [], []
    for _ in range(args.reps):
        model.load_state_dict(state)           # BN running statistics back to the start as well
        model.zero_grad(set_to_none=True)
        out = model(imgs)
        ce, focal, dice, fs = seg_objective(out, pngs, labels, cls_w, 5)
        (focal + dice).backward()
        g2 = torch.zeros((), device="cuda", dtype=torch.float64)
        for p in model.parameters():
            if p.grad is not None:
                g2 += p.grad.double().pow(2).sum()
        focals.append(float(focal)); dices.append(float(dice)); gnorms.append(float(g2.sqrt()))
    sp = lambda v: (max(v) - min(v)) / (abs(sum(v) / len(v)) + 1e-30)
    print("%-5s %4d^2 B=%-3d focal %s  spread %.2e | dice spread %.2e | |grad| spread %.2e" %
          (dt, size, bsz, " ".join("%.5f" % f for f in focals), sp(focals), sp(dices), sp(gnorms)), flush=True)
    del model
    torch.cuda.empty_cache()
