"""Kernel-time breakdown of one training step with torch.profiler (CUPTI), GPU box only.
    python tools/profile_step.py [--batch 32] [--steps 2]
Prints the top kernels by total device time (name, calls, total ms, % of step)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

from cervix_b200.engine import SegTrainer
from cervix_b200.nets.deeplabv3_plus import DeepLab
from bench import synthetic_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--backbone", default="xception")
args = ap.parse_args()

torch.manual_seed(0)
model = DeepLab(5, args.backbone, False, 16).set_compute_dtype(torch.bfloat16).cuda().train()
trainer = SegTrainer(model, cls_weights=[1, 1, 5, 3, 4])
imgs, pngs, labels = [t.cuda() for t in synthetic_batch(args.batch, 512, seed=0)]
for _ in range(2):
    trainer.step(imgs, pngs, labels)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(args.steps):
        trainer.step(imgs, pngs, labels)
    torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / args.steps * 1e3
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        r = rows.setdefault(ev.name, [0, 0.0])
        r[0] += 1
        r[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
tot = sum(r[1] for r in rows.values())
print("batch %d: wall %.1f ms/step (under profiler), sum of kernel time %.1f ms/step, %d distinct kernels" %
      (args.batch, wall, tot / args.steps / 1e3, len(rows)))
for name, (calls, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%7.2f ms %5.1f%% %6d calls  %s" % (us / args.steps / 1e3, 100 * us / tot, calls // args.steps, name[:110]))
