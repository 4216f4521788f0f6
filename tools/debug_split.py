import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cervix_b200.engine import SegTrainer, GradGather
from cervix_b200.nets.deeplabv3_plus import DeepLab
from cervix_b200 import ops

torch.manual_seed(2)
model = DeepLab(5, "xception", False, 16).set_compute_dtype(torch.float32).cuda().train()
for m in model.modules():
    if isinstance(m, torch.nn.Dropout):
        m.p = 0.0
g = torch.Generator().manual_seed(3)
pngs = torch.randint(0, 6, (4, 64, 64), generator=g).cuda()
imgs = torch.rand(4, 3, 64, 64, generator=g).cuda()
tr = SegTrainer(model, lr=0.0, optimizer="sgd", cls_weights=[1, 1, 5, 3, 4])
names = [n for n, _ in model.named_parameters()]
tr._forward_backward(imgs, pngs, None)
g1 = tr.flat.grad.clone()
tr._forward_backward(imgs, pngs, None)
g1b = tr.flat.grad.clone()
print("single vs single: max diff %.3e (max |g| %.3e)" % (float((g1 - g1b).abs().max()), float(g1.abs().max())))
early, late = tr._split_plan()
tr.flat.detach_grads()
tr._phase_a(imgs, pngs, None, late)
late_grads = tr._split_late_grads
early_grads = tr._phase_b(early)
g2 = torch.zeros_like(g1)
for idx, grads in ((late, late_grads), (early, early_grads)):
    for i, gr in zip(idx, grads):
        if gr is not None:
            o = tr.flat.offsets[i]
            g2[o:o + gr.numel()] = gr.reshape(-1)
rows = []
for i, n in enumerate(names):
    o = tr.flat.offsets[i]; k = tr.flat.params[i].numel()
    a, b = g1[o:o + k], g2[o:o + k]
    rows.append((float((a - b).abs().max()) / (float(a.abs().max()) + 1e-30), n, float(a.abs().max())))
rows.sort(reverse=True)
for r in rows[:25]:
    print("%.3e  %-50s |g|max %.3e" % r)
print("early count", len(early), "first late", names[late[0]])

def report(tag, ga, gb):
    rows = []
    for i, n in enumerate(names):
        o = tr.flat.offsets[i]; k = tr.flat.params[i].numel()
        a, b = ga[o:o + k], gb[o:o + k]
        rows.append((float((a - b).abs().max()) / (float(a.abs().max()) + 1e-30), n, float(a.abs().max())))
    bad = [r for r in rows if r[0] > 1e-4]
    print("== %s: %d / %d parameters differ by more than 1e-4 relative" % (tag, len(bad), len(rows)))
    for r in rows[:80]:
        if r[0] > 1e-4:
            print("   %.3e  %-50s |g|max %.3e" % r)

from cervix_b200.backend import get_backend
Bk = get_backend()
for use_arena in (True, False):
    if not use_arena:
        Bk.zero_arena_begin = None
    torch.manual_seed(2)
    m2 = DeepLab(5, "xception", False, 16).set_compute_dtype(torch.float32).cuda().train()
    m2.load_state_dict(model.state_dict())
    for m in m2.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    t2 = SegTrainer(m2, lr=0.0, optimizer="sgd", cls_weights=[1, 1, 5, 3, 4])
    assert t2.capture_split(imgs, pngs, None, warmup=1) is not None
    for s in range(2):
        t2.step_graphed(imgs, pngs)
        torch.cuda.synchronize()
        report("graph split arena=%s step %d" % (use_arena, s), g1, t2.flat.grad)
