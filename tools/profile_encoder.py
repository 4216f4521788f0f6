"""Kernel-time breakdown of the ResNet-101 patch encoder forward (256 patches of 256x256, bf16, eval), GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

from cervix_b200.multimodal.patch_encoder import ResNet101Encoder, split_patches

torch.manual_seed(0)
enc = ResNet101Encoder().cuda().eval()
imgs = torch.rand(16, 3, 512, 512, device="cuda")
with torch.no_grad():
    for _ in range(2):
        enc(split_patches(imgs))
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        enc(split_patches(imgs))
        torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        r = rows.setdefault(ev.name, [0, 0.0])
        r[0] += 1
        r[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
tot = sum(r[1] for r in rows.values())
print("256 patches: sum of kernel time %.2f ms, %d distinct kernels" % (tot / 1e3, len(rows)))
for name, (calls, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:20]:
    print("%7.2f ms %5.1f%% %6d calls  %s" % (us / 1e3, 100 * us / tot, calls, name[:110]))
