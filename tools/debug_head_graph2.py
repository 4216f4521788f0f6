import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from cervix_b200.engine import FusionTrainer
from cervix_b200.multimodal.my_mae_model import fusion_model_mae_2, get_edge_index_full, get_edge_index_image
from tests.test_fusion_gpu import rnd

def run(types, order):
    types = list(types)
    T, G = len(types), 6
    all_edges = {"imgN": get_edge_index_image(), "imgA": get_edge_index_image(), "imgL": get_edge_index_image(), "cli": get_edge_index_full(4)}
    edges = {m: all_edges[m] for m in types}
    rng = np.random.RandomState(1)
    def batch(seed):
        feats = {m: rnd(G, 4 if m == "cli" else 16, 1024, seed=seed * 10 + i) for i, m in enumerate(types)}
        labels = torch.from_numpy(rng.randint(0, 4, G)).cuda()
        masks = np.ones((G, T), dtype=bool)
        masks[np.arange(G), rng.randint(0, T, G)] = False
        return feats, labels, masks
    batches = [batch(s) for s in range(4)]
    trainers = []
    for _ in range(2):
        torch.manual_seed(0)
        head = fusion_model_mae_2(1024, 512, 512, 0.3, T).cuda().eval()
        trainers.append(FusionTrainer(head, types, lr=1e-3, weight_decay=1e-3))
    eager, graphed = trainers if order == 0 else trainers[::-1]
    f0, l0, m0 = batches[0]
    names = [n for n, _ in eager.head.named_parameters()]
    def report(tag):
        d = (eager.flat.data - graphed.flat.data).abs(); g = (eager.flat.grad - graphed.flat.grad).abs()
        worst = max(((float(d[o:o + p.numel()].max()), n) for n, p, o in zip(names, eager.flat.params, eager.flat.offsets)))
        print("  %s: max |dparam| %.3e (%s) max |dgrad| %.3e" % (tag, float(d.max()), worst[1], float(g.max())))
    report("init")
    graphed.capture(f0, edges, l0, m0, warmup=2)
    [float(eager.step(f0, edges, l0, m0)) for _ in range(2)]
    report("after warm-up")
    for k, (f, l, m) in enumerate(batches[1:]):
        le = float(eager.step(f, edges, l, m)); lg = float(graphed.step_graphed(f, l, m))
        print(" step", k, le, lg); report("")
for order in (0, 1):
    print("ORDER", order)
    run(("imgN", "imgA", "imgL", "cli"), order)
