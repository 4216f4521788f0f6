"""Debug: every torch.empty / empty_like is filled with NaN (floats) or a large value (ints); a NaN or a crash in the head's
loss / gradients then points at a kernel that reads memory nobody wrote (GPU box only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

_empty, _empty_like = torch.empty, torch.empty_like
def _poison(t):
    if t.is_cuda and t.numel():
        if t.dtype.is_floating_point:
            t.fill_(float("nan"))
        elif t.dtype in (torch.int32, torch.int64, torch.uint8):
            t.fill_(113)
    return t
torch.empty = lambda *a, **k: _poison(_empty(*a, **k))
torch.empty_like = lambda *a, **k: _poison(_empty_like(*a, **k))

from cervix_b200.engine import FusionTrainer
from cervix_b200.multimodal.my_mae_model import fusion_model_mae_2, fusion_objective, get_edge_index_full, get_edge_index_image

for types in (["imgN", "imgA", "imgL", "cli"], ["imgN", "imgL"], ["imgN", "imgA", "imgL"]):
    T, G = len(types), 6
    all_edges = {"imgN": get_edge_index_image(), "imgA": get_edge_index_image(), "imgL": get_edge_index_image(), "cli": get_edge_index_full(4)}
    edges = {m: all_edges[m] for m in types}
    rng = np.random.RandomState(1)
    feats = {m: torch.randn(G, 4 if m == "cli" else 16, 1024).cuda() for m in types}
    labels = torch.from_numpy(rng.randint(0, 4, G)).cuda()
    masks = np.ones((G, T), dtype=bool); masks[np.arange(G), rng.randint(0, T, G)] = False
    torch.manual_seed(0)
    for mode in ("eval", "train"):
        head = fusion_model_mae_2(1024, 512, 512, 0.3, T).cuda()
        head.train(mode == "train")
        out = head.forward_batch(feats, edges, types, types, masks, True)
        for k, v in out.items():
            if torch.is_tensor(v) and v.dtype.is_floating_point and not bool(torch.isfinite(v).all()):
                print("NON-FINITE forward output", types, mode, k)
        loss = fusion_objective(out, labels, masks)
        loss.backward()
        bad = [n for n, p in head.named_parameters() if p.grad is not None and not bool(torch.isfinite(p.grad).all())]
        print(types, mode, "loss", float(loss), "non-finite grads:", bad[:8])
# the segmentation path too: one bf16 train step of a small model
from cervix_b200.nets.deeplabv3_plus import DeepLab
from cervix_b200.nets.deeplabv3_training import seg_objective
for bb, dtype in (("xception", torch.bfloat16), ("mobilenet", torch.bfloat16), ("xception", torch.float32)):
    torch.manual_seed(0)
    model = DeepLab(5, bb, False, 16).set_compute_dtype(dtype).cuda().train()
    imgs = torch.rand(4, 3, 96, 96).cuda(); pngs = torch.randint(0, 6, (4, 96, 96)).cuda()
    ce, focal, dice, fs = seg_objective(model(imgs), pngs, None, torch.tensor([1., 1, 5, 3, 4]).cuda(), 5)
    (focal + dice).backward()
    bad = [n for n, p in model.named_parameters() if p.grad is not None and not bool(torch.isfinite(p.grad).all())]
    print(bb, dtype, "losses", float(ce), float(focal), float(dice), float(fs), "non-finite grads:", bad[:8])
