"""Debug: eager vs eager vs graphed FusionTrainer trajectories, step by step (GPU box only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from cervix_b200.engine import FusionTrainer
from cervix_b200.multimodal.my_mae_model import fusion_model_mae_2, get_edge_index_full, get_edge_index_image

types = ["imgN", "imgA", "imgL", "cli"]
T, G = 4, 6
edges = {"imgN": get_edge_index_image(), "imgA": get_edge_index_image(), "imgL": get_edge_index_image(), "cli": get_edge_index_full(4)}
for trial in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    rng = np.random.RandomState(1)
    def rnd(*shape, seed=0):
        return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).cuda()
    def batch(seed):
        feats = {m: rnd(G, 4 if m == "cli" else 16, 1024, seed=seed * 10 + i) for i, m in enumerate(types)}
        labels = torch.from_numpy(rng.randint(0, 4, G)).cuda()
        masks = np.ones((G, T), dtype=bool)
        masks[np.arange(G), rng.randint(0, T, G)] = False
        return feats, labels, masks
    batches = [batch(s) for s in range(4)]
    trs = []
    for _ in range(3):
        torch.manual_seed(0)
        head = fusion_model_mae_2(1024, 512, 512, 0.3, T).cuda().eval()
        trs.append(FusionTrainer(head, types, lr=1e-3, weight_decay=1e-3))
    A, Bt, Cg = trs
    f0, l0, m0 = batches[0]
    Cg.capture(f0, edges, l0, m0, warmup=2)
    for _ in range(2):
        A.step(f0, edges, l0, m0); Bt.step(f0, edges, l0, m0)
    names = [n for n, _ in A.head.named_parameters()]
    def report(tag, X, Y):
        d = (X.flat.data - Y.flat.data).abs()
        g = (X.flat.grad - Y.flat.grad).abs()
        worst = max(((float(d[o:o + p.numel()].max()), n) for n, p, o in zip(names, X.flat.params, X.flat.offsets)))
        print("  %s: max |dparam| %.3e (%s) max |dgrad| %.3e (|grad| max %.3e)" % (tag, float(d.max()), worst[1], float(g.max()), float(X.flat.grad.abs().max())))
    print("trial", trial, "after warm-up:"); report("A-B", A, Bt); report("A-C", A, Cg)
    for k, (f, l, m) in enumerate(batches[1:]):
        la = float(A.step(f, edges, l, m)); lb = float(Bt.step(f, edges, l, m)); lc = float(Cg.step_graphed(f, l, m))
        print(" step", k, "loss A %.7f B %.7f C %.7f" % (la, lb, lc)); report("A-B", A, Bt); report("A-C", A, Cg)
