import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from cervix_b200.engine import SegTrainer
from test_engine_gpu import _small_model, _batches
imgs, pngs, _ = _batches(1, bsz=4, size=64, seed=3)[0]
imgs, pngs = imgs.cuda(), pngs.cuda()
kw = dict(lr=1e-2, optimizer="sgd", momentum=0.9, weight_decay=1e-4, cls_weights=[1, 1, 5, 3, 4])
a = SegTrainer(_small_model(torch.float32, seed=2, bb="xception"), **kw).capture(imgs, pngs, None, warmup=1)
b = SegTrainer(_small_model(torch.float32, seed=2, bb="xception"), **kw).capture(imgs, pngs, None, warmup=1)
c = SegTrainer(_small_model(torch.float32, seed=2, bb="xception"), **kw).capture_split(imgs, pngs, None, warmup=1)
print("weights after warm-up: a-b %.3e a-c %.3e" % (float((a.flat.data - b.flat.data).abs().max()), float((a.flat.data - c.flat.data).abs().max())))
for s in range(3):
    a.step_graphed(imgs, pngs); b.step_graphed(imgs, pngs); c.step_graphed(imgs, pngs)
    ga, gb, gc = a.flat.grad, b.flat.grad, c.flat.grad
    print("step %d  one-one %.3e  one-split %.3e  max|g| %.3e   w: a-b %.3e a-c %.3e" % (
        s, float((ga - gb).abs().max()), float((ga - gc).abs().max()), float(ga.abs().max()),
        float((a.flat.data - b.flat.data).abs().max()), float((a.flat.data - c.flat.data).abs().max())))
