"""CPU tests of the training-step engine (flat parameters, fused optimizer call, bucketed
gradient all-reduce) with the emulation backend; the N>1 path runs as a world_size-2 gloo job."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cervix_b200.backend as backend
from cervix_b200.engine import SegTrainer
from cervix_b200.nets.deeplabv3_plus import DeepLab
from oracle import deeplab_ref as O
from oracle import losses_ref as L
from tests.emu_backend import EmuBackend

CLS_W = [1, 1, 5, 3, 4]


def _model(seed=3):
    m = DeepLab(5, "mobilenet", False, 16).set_compute_dtype(torch.float32)
    m.load_state_dict(O.make_state("mobilenet", 5, 16, seed=seed, randomize_bn_stats=False))
    m.train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m


def _oracle_grads(state, imgs, pngs, labels):
    st = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
          for k, v in state.items()}
    y = O.deeplab_forward(imgs, st, "mobilenet", 16, True)
    w = torch.tensor(CLS_W, dtype=torch.float32)
    (L.focal_loss(y, pngs, w, 5) + L.dice_loss(y, labels)).backward()
    return {k: v.grad for k, v in st.items() if v.dtype.is_floating_point and v.grad is not None}


def test_single_rank_step_matches_torch_adam():
    prev = backend.set_backend(EmuBackend())
    try:
        model = _model()
        state0 = {k: v.clone() for k, v in model.state_dict().items()}
        trainer = SegTrainer(model, lr=1e-3, cls_weights=CLS_W)
        assert trainer.flat.data.numel() >= sum(p.numel() for p in model.parameters())
        imgs, pngs, labels = O.synthetic_batch(4, 64, seed=1)
        res = trainer.step(imgs, pngs, labels)
        assert res.shape == (4,) and torch.isfinite(res).all()
        grads = _oracle_grads(state0, imgs, pngs, labels)
        # one Adam step from zero state moves every weight by ~lr*sign(grad): check direction on a big tensor
        k = "cat_conv.4.weight"
        delta = dict(model.named_parameters())[k].detach() - state0[k]
        big = grads[k].abs() > 0.2 * grads[k].abs().max()
        assert torch.equal(torch.sign(delta[big]), -torch.sign(grads[k][big]))
        assert abs(float(delta[big].abs().mean()) - 1e-3) < 1e-4
        # parameters are views of the flat buffer, gradients too
        p0 = trainer.flat.params[0]
        assert p0.data_ptr() == trainer.flat.data.data_ptr() and p0.grad.data_ptr() == trainer.flat.grad.data_ptr()
    finally:
        backend.set_backend(prev)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    backend.set_backend(EmuBackend())
    torch.set_num_threads(2)
    model = _model()
    trainer = SegTrainer(model, lr=1e-3, cls_weights=CLS_W, world_size=world, bucket_mb=2.0)
    assert len(trainer.buckets) >= 3
    imgs, pngs, labels = O.synthetic_batch(2, 64, seed=10 + rank)
    trainer.step(imgs, pngs, labels)
    torch.save({"grad": trainer.flat.grad.clone(), "data": trainer.flat.data.clone()}, os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_manual_average(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    # all ranks hold the same summed gradient and took the same optimizer step
    assert torch.equal(r0["grad"], r1["grad"]) and torch.equal(r0["data"], r1["data"])
    # and it equals the sum of the two shards' single-process gradients (per-rank BatchNorm statistics)
    prev = backend.set_backend(EmuBackend())
    try:
        total = None
        for rank in range(2):
            model = _model()
            tr = SegTrainer(model, lr=0.0, cls_weights=CLS_W)
            imgs, pngs, labels = O.synthetic_batch(2, 64, seed=10 + rank)
            tr.step(imgs, pngs, labels)
            total = tr.flat.grad.clone() if total is None else total + tr.flat.grad
    finally:
        backend.set_backend(prev)
    assert float((r0["grad"] - total).abs().max()) <= 1e-5 * float(total.abs().max())


# ---- fusion head (classifier) data parallelism: patients sharded across ranks ----------------
def _fusion_setup(rank, types):
    import numpy as np
    from cervix_b200.multimodal.my_mae_model import fusion_model_mae_2
    from oracle import fusion_ref as FR
    from tests.fusion_cases import batch_of
    torch.manual_seed(0)
    head = fusion_model_mae_2(1024, 512, 512, 0.3, len(types))
    head.load_state_dict(FR.randomize_state(head.state_dict(), seed=3), strict=True)
    head.eval()                                       # dropout off: the two runs must see the same function
    patients = [FR.synthetic_patient(2 * rank + i) for i in range(2)]
    feats, edges = batch_of(patients, types, "cpu")
    masks = np.ones((2, len(types)), dtype=bool)
    masks[0, rank % len(types)] = False
    masks[1, (rank + 1) % len(types)] = False
    labels = torch.tensor([rank % 4, (rank + 2) % 4])
    return head, feats, edges, labels, masks


def _fusion_worker(rank, world, port, out_dir):
    from cervix_b200.engine import FusionTrainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    backend.set_backend(EmuBackend())
    torch.set_num_threads(2)
    types = ["imgN", "imgA", "imgL", "cli"]
    head, feats, edges, labels, masks = _fusion_setup(rank, types)
    tr = FusionTrainer(head, types, lr=1e-3, world_size=world)
    loss = tr.step(feats, edges, labels, masks)
    torch.save({"grad": tr.flat.grad.clone(), "data": tr.flat.data.clone(), "loss": float(loss)},
               os.path.join(out_dir, "f%d.pt" % rank))
    dist.destroy_process_group()


def test_fusion_two_rank_gloo_allreduce_matches_manual_sum(tmp_path):
    from cervix_b200.engine import FusionTrainer
    port = 31500 + os.getpid() % 2000
    mp.spawn(_fusion_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "f0.pt"), torch.load(tmp_path / "f1.pt")
    assert torch.equal(r0["grad"], r1["grad"]) and torch.equal(r0["data"], r1["data"])
    assert r0["loss"] != r1["loss"]                       # each rank worked on its own patients
    prev = backend.set_backend(EmuBackend())
    try:
        types = ["imgN", "imgA", "imgL", "cli"]
        total, start = None, None
        for rank in range(2):
            head, feats, edges, labels, masks = _fusion_setup(rank, types)
            tr = FusionTrainer(head, types, lr=0.0, weight_decay=0.0)
            start = tr.flat.data.clone()
            tr.step(feats, edges, labels, masks)
            total = tr.flat.grad.clone() if total is None else total + tr.flat.grad
        # the optimizer saw the mean of the two shards' gradients: redo Adam's first step by hand
        g = total / 2 + 5e-4 * start
        want = start - 1e-3 * g / (g.abs() + 1e-8)
    finally:
        backend.set_backend(prev)
    assert float((r0["grad"] - total).abs().max()) <= 1e-5 * float(total.abs().max())
    assert float((r0["data"] - want).abs().max()) <= 2e-6


def test_split_backward_plan_and_eager_phases_match_the_single_backward():
    """Host logic of the two-graph data-parallel step (SegTrainer.capture_split) without a GPU: the plan puts Xception's
    entry flow - and nothing else - before the cut, declines models without a cut or with frozen parameters, and the two
    phases run eagerly (forward + late backward through the cut proxies, then the entry-flow backward from the gradients
    left at the cuts) reproduce the gradients of one ordinary backward on the emulated ABI."""
    prev = backend.set_backend(EmuBackend())
    try:
        torch.manual_seed(1)
        model = DeepLab(5, "xception", False, 16).set_compute_dtype(torch.float32).train()
        for mod in model.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        tr = SegTrainer(model, lr=0.0, optimizer="sgd", cls_weights=CLS_W)
        names = [n for n, _ in model.named_parameters()]
        early, late = tr._split_plan()
        assert early == list(range(len(early))) and late == list(range(len(early), len(names)))
        assert all(names[i].split(".")[1] in SegTrainer._ENTRY_MODULES for i in early)
        assert names[late[0]].startswith("backbone.block4.") and not any(n.startswith("backbone.block3.") for n in names[late[0]:])
        assert sum(tr.flat.params[i].numel() for i in early) < 0.1 * tr.flat.numel
        g = torch.Generator().manual_seed(2)
        imgs, pngs = torch.rand(2, 3, 32, 32, generator=g), torch.randint(0, 6, (2, 32, 32), generator=g)
        tr._forward_backward(imgs, pngs, None)
        want = tr.flat.grad.clone()
        tr.flat.detach_grads()
        tr._phase_a(imgs, pngs, None, late)
        late_grads = tr._split_late_grads
        early_grads = tr._phase_b(early)
        got = torch.zeros_like(want)
        for idx, grads in ((late, late_grads), (early, early_grads)):
            for i, gr in zip(idx, grads):
                if gr is not None:
                    o = tr.flat.offsets[i]
                    got[o:o + gr.numel()] = gr.reshape(-1)
        assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())
        # a frozen parameter or a backbone without a cut: no plan (the caller keeps the one-graph step)
        next(model.backbone.parameters()).requires_grad_(False)
        assert SegTrainer(model, lr=0.0, optimizer="sgd", cls_weights=CLS_W)._split_plan() is None
        assert SegTrainer(_model(), lr=0.0, optimizer="sgd", cls_weights=CLS_W)._split_plan() is None
    finally:
        backend.set_backend(prev)
