"""Kernel-time breakdown of the fusion head's graph-captured train step (GPU box only).
    python tools/profile_head.py [--patients 16]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from cervix_b200.engine import FusionTrainer
from cervix_b200.multimodal.my_mae_model import fusion_model_mae_2, get_edge_index_full, get_edge_index_image

ap = argparse.ArgumentParser()
ap.add_argument("--patients", type=int, default=16)
args = ap.parse_args()
G = args.patients
types = ["imgN", "imgA", "imgL", "cli"]
torch.manual_seed(0)
head = fusion_model_mae_2(1024, 512, 512, 0.3, 4).cuda().train()
tr = FusionTrainer(head, types)
edges = {"imgN": get_edge_index_image(), "imgA": get_edge_index_image(), "imgL": get_edge_index_image(),
         "cli": get_edge_index_full(4)}
feats = {m: torch.randn(G, 4 if m == "cli" else 16, 1024, device="cuda") for m in types}
labels = torch.randint(0, 4, (G,), device="cuda")
masks = np.ones((G, 4), dtype=bool); masks[np.arange(G), np.arange(G) % 4] = False
tr.capture(feats, edges, labels, masks)
for _ in range(3):
    tr.step_graphed(feats, labels, masks)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    tr.step_graphed(feats, labels, masks)
e1.record(); torch.cuda.synchronize()
print("graph replay: %.2f ms/step" % (e0.elapsed_time(e1) / 10))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        tr.step_graphed(feats, labels, masks)
    torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        r = rows.setdefault(ev.name, [0, 0.0]); r[0] += 1; r[1] += ev.device_time_total
tot = sum(r[1] for r in rows.values())
print("sum of kernel time %.2f ms/step, %d launches/step" % (tot / 2e3, sum(r[0] for r in rows.values()) // 2))
for name, (calls, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%7.3f ms %5.1f%% %5d calls  %s" % (us / 2e3, 100 * us / tot, calls // 2, name[:120]))
