"""Per-backend-call time of one eager training step, with the tensor shapes of each call (GPU box only).
    python tools/profile_ops.py [--batch 32] [--top 70]
Every CudaBackend method is wrapped with CUDA events and a synchronize, so small calls are inflated by the sync but the
large ones (the ones worth looking at) are accurate.  Output: calls grouped by (method, shapes), sorted by total time,
with the bytes of their tensor arguments and results and the implied GB/s."""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from cervix_b200.backend import get_backend
from cervix_b200.engine import SegTrainer
from cervix_b200.nets.deeplabv3_plus import DeepLab
from bench import synthetic_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--top", type=int, default=70)
args = ap.parse_args()

torch.manual_seed(0)
model = DeepLab(5, "xception", False, 16).set_compute_dtype(torch.bfloat16).cuda().train()
trainer = SegTrainer(model, cls_weights=[1, 1, 5, 3, 4])
imgs, pngs, labels = [t.cuda() for t in synthetic_batch(args.batch, 512, seed=0)]
for _ in range(2):
    trainer.step(imgs, pngs, None)
torch.cuda.synchronize()

B = get_backend()
log = []
depth = [0]


def tensors(obj):
    if torch.is_tensor(obj):
        yield obj
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            yield from tensors(o)


def wrap(name, fn):
    def inner(*a, **k):
        if depth[0]:
            return fn(*a, **k)
        depth[0] += 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        try:
            out = fn(*a, **k)
        finally:
            depth[0] -= 1
        e1.record()
        torch.cuda.synchronize()
        ins = [t for t in tensors(a)] + [t for t in tensors(list(k.values()))]
        outs = list(tensors(out))
        geom = next((x for x in a if hasattr(x, "cin")), None)
        key = (name, tuple(tuple(t.shape) for t in ins if t.numel() > 4096),
               None if geom is None else (geom.kh, geom.stride, geom.dil))
        nbytes = sum(t.numel() * t.element_size() for t in ins + outs)
        log.append((key, e0.elapsed_time(e1), nbytes))
        return out
    return inner


for name in dir(B):
    if name.startswith("_") or name in ("lib", "name", "is_sm100"):
        continue
    fn = getattr(B, name)
    if callable(fn):
        setattr(B, name, wrap(name, fn))

trainer.step(imgs, pngs, None)
torch.cuda.synchronize()
agg = collections.OrderedDict()
for key, ms, nbytes in log:
    r = agg.setdefault(key, [0, 0.0, 0])
    r[0] += 1; r[1] += ms; r[2] += nbytes
tot = sum(r[1] for r in agg.values())
print("batch %d: %d backend calls, %.1f ms summed (eager, sync per call)" % (args.batch, len(log), tot))
by_method = collections.Counter()
for (name, _, _), (c, ms, _) in agg.items():
    by_method[name] += ms
print("by method: " + ", ".join("%s %.2f" % kv for kv in by_method.most_common(30)))
for key, (calls, ms, nbytes) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
    print("%7.3f ms %4d x %7.1f us %7.0f MB/call %6.0f GB/s  %s %s %s" % (
        ms, calls, ms / calls * 1e3, nbytes / calls / 1e6, nbytes / (ms * 1e-3) / 1e9, key[0], key[2] or "",
        " ".join("x".join(map(str, s)) for s in key[1])))
