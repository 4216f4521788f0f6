"""GPU tests of the step engine's host<->device plumbing."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_batch_prefetcher_double_buffers_keep_batches_intact():
    """Six distinct pinned host batches through the two alternating device buffer sets, with a consumer that is slow
    enough for the next copy to be issued while the previous batch is still being read."""
    from cervix_b200.engine import BatchPrefetcher
    torch.manual_seed(0)
    host = [(torch.full((4, 3, 64, 64), float(i)).pin_memory(), torch.full((4, 64, 64), i, dtype=torch.int64).pin_memory(), None)
            for i in range(6)]
    big = torch.randn(4096, 4096, device="cuda")
    sums, seen = [], 0
    for i, (a, b, c) in enumerate(BatchPrefetcher(iter(host))):
        assert c is None and a.is_cuda and b.is_cuda
        for _ in range(3):
            big = (big @ big).clamp_(-1, 1)          # keeps the compute stream busy past the next prefetch
        sums.append((a.sum() + 0 * big[0, 0], b.sum()))
        seen += 1
    assert seen == 6
    for i, (sa, sb) in enumerate(sums):
        assert float(sa) == i * 4 * 3 * 64 * 64 and int(sb) == i * 4 * 64 * 64


def test_uint8_batches_match_the_float_contract():
    """The decoded uint8 [B,H,W,3] pixels + uint8 class map (what the reference's loader holds before its float
    conversion, dataloader.py:40-42) against the fp32 NCHW / int64 tensors the loader hands to fit_one_epoch."""
    import numpy as np
    from cervix_b200.engine import SegTrainer
    from cervix_b200.nets.deeplabv3_plus import DeepLab
    rng = np.random.RandomState(0)
    u8 = torch.from_numpy(rng.randint(0, 256, (2, 64, 64, 3), dtype=np.uint8)).cuda()
    lab = rng.randint(0, 5, (2, 64, 64)).astype(np.uint8)
    lab[0, :4] = 255                                          # VOC-style white border -> ignore label
    lab_u8 = torch.from_numpy(lab).cuda()
    f32 = (u8.float() / 255.0).permute(0, 3, 1, 2).contiguous()
    i64 = lab_u8.long().clamp(max=5)
    torch.manual_seed(0)
    model = DeepLab(5, "mobilenet", False, 16).set_compute_dtype(torch.bfloat16).cuda().eval()
    with torch.no_grad():
        a, b = model(u8), model(f32)
    assert a.shape == b.shape == (2, 5, 64, 64)
    assert torch.equal(a, b)                                  # same bf16 activations enter the stem
    # gradients through the trainer, BatchNorm on its running statistics: with batch statistics this tiny train-mode
    # network (2 images, 4x4 high-level features) amplifies the summation order of the fp32 atomics so much that two
    # steps on IDENTICAL inputs give gradients with a cosine of ~0.3, which would hide the input path
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
    tr = SegTrainer(model, lr=0.0, cls_weights=[1, 1, 5, 3, 4])
    la = tr.step(u8, lab_u8).clone()
    ga = tr.flat.grad.clone()
    lb = tr.step(f32, i64).clone()
    gb = tr.flat.grad
    assert float((la - lb).abs().max()) <= 1e-4 * float(lb.abs().max())
    cos = float(torch.dot(ga, gb) / (ga.norm() * gb.norm()))
    assert cos > 0.999, cos


# ----------------------------------------------------------------------------------------------------------------------
# SegTrainer as the engine behind fit_one_epoch (VERDICT r01 items 6, 9 and the advisor's engine findings)
def _small_model(dtype=torch.float32, seed=0, bb="mobilenet"):
    from cervix_b200.nets.deeplabv3_plus import DeepLab
    torch.manual_seed(seed)
    model = DeepLab(5, bb, False, 16).set_compute_dtype(dtype).cuda().train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return model


def _batches(n, bsz=4, size=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        pngs = torch.randint(0, 6, (bsz, size, size), generator=g)
        out.append((torch.rand(bsz, 3, size, size, generator=g), pngs, torch.eye(6)[pngs]))
    return out


class _Hist:
    def __init__(self):
        self.losses, self.val_loss = [], []

    def append_loss(self, e, a, b):
        self.losses.append(a); self.val_loss.append(b)


class _Ev:
    def on_epoch_end(self, e, m):
        self.seen = e


@pytest.mark.parametrize("kind", ["adam", "sgd"])
def test_fit_one_epoch_fast_engine_equals_the_eager_loop(tmp_path, monkeypatch, kind):
    """The reference's entry point on the fast engine (flat parameters, fused optimizer, CUDA graph from the third batch,
    hyper-parameters re-read from the torch optimizer every epoch) against the plain autograd + ``optimizer.step()`` loop
    of the same function: same batches, fp32 engine, the per-epoch mean losses of both must agree, across a
    learning-rate change between the epochs (``set_optimizer_lr``, train.py:575)."""
    import numpy as np
    from cervix_b200.utils.utils_fit import fit_one_epoch
    cls_w = np.array([1, 1, 5, 3, 4], np.float32)
    data = _batches(6)
    results = {}
    for mode in ("fast", "eager"):
        monkeypatch.setenv("CERVIX_FIT_EAGER", "1" if mode == "eager" else "0")
        model = _small_model()
        if kind == "adam":
            opt = torch.optim.Adam(model.parameters(), 3e-4, betas=(0.9, 0.999), weight_decay=0)
        else:
            opt = torch.optim.SGD(model.parameters(), 5e-3, momentum=0.9, nesterov=True, weight_decay=1e-4)
        hist, ev = _Hist(), _Ev()
        save = tmp_path / (mode + kind)
        save.mkdir()
        for epoch in range(2):
            for gparam in opt.param_groups:
                gparam["lr"] = (3e-4 if kind == "adam" else 5e-3) * (1.0 if epoch == 0 else 0.3)
            fit_one_epoch(model, model, hist, ev, opt, epoch, 6, 1, data, data[:1], 2, True, True, True, cls_w, 5, False, None,
                          1, str(save))
        results[mode] = (hist.losses, hist.val_loss, {k: v.clone() for k, v in model.state_dict().items()})
        if mode == "fast":
            tr = model._cvx_trainer
            assert tr.graph is not None and tr.t == 12          # the graph was captured and every batch was a step
            assert abs(tr.lr - opt.param_groups[0]["lr"]) < 1e-12
            assert (save / "last_epoch_trainer_state.pth").exists()
            assert len(opt.state) == 0                          # the torch optimizer was never stepped
    (lf, vf, sf), (le, ve, se) = results["fast"], results["eager"]
    # batch-statistics BatchNorm at this size amplifies the summation order of the weight-gradient atomics from step to
    # step, so two 12-step trajectories agree to per cent, not to rounding; the exact one-step
    # comparison against torch.optim is test_trainer_step_equals_torch_optim_step below
    assert np.allclose(lf, le, rtol=3e-2), (lf, le)
    assert np.allclose(vf, ve, rtol=6e-2), (vf, ve)
    assert lf[1] < lf[0]


@pytest.mark.parametrize("kind", ["adam", "sgd"])
def test_trainer_step_equals_torch_optim_step(kind):
    """One SegTrainer step against autograd + ``torch.optim`` from the same weights on the same batch (fp32 engine): the
    fused optimizer launch applies torch's update rule (Adam with bias correction / SGD with nesterov momentum and
    weight decay), first step and second step (momentum buffers in use)."""
    from cervix_b200.engine import SegTrainer
    from cervix_b200.nets.deeplabv3_training import seg_objective
    imgs, pngs, _ = _batches(1)[0]
    imgs, pngs = imgs.cuda(), pngs.cuda()
    cls_w = torch.tensor([1, 1, 5, 3, 4], dtype=torch.float32, device="cuda")
    ma, mb = _small_model(seed=1), _small_model(seed=1)
    if kind == "adam":
        opt = torch.optim.Adam(mb.parameters(), 1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-3)
        tr = SegTrainer(ma, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-3, cls_weights=cls_w)
    else:
        opt = torch.optim.SGD(mb.parameters(), 1e-2, momentum=0.9, nesterov=True, weight_decay=1e-4)
        tr = SegTrainer(ma, lr=1e-2, optimizer="sgd", momentum=0.9, nesterov=True, weight_decay=1e-4, cls_weights=cls_w)
    for step in range(2):
        before = tr.flat.data.clone()
        mb.load_state_dict(ma.state_dict())           # same weights and BatchNorm buffers going into the step
        tr.step(imgs, pngs)
        opt.zero_grad()
        _, focal, dice, _ = seg_objective(mb(imgs), pngs, None, cls_w, 5)
        (focal + dice).backward()
        opt.step()
        got = torch.cat([p.detach().reshape(-1) for p in ma.parameters()])
        want = torch.cat([p.detach().reshape(-1) for p in mb.parameters()])
        grad = torch.cat([p.grad.reshape(-1) for p in mb.parameters()])
        if kind == "adam":   # Adam turns rounding-level gradients into +-lr: compare where the gradient is a real number
            sel = grad.abs() > 1e-4 * grad.abs().max()
            assert float(sel.float().mean()) > 0.3
        else:
            sel = torch.ones_like(grad, dtype=torch.bool)
        step_size = float((want - torch.cat([b.reshape(-1) for b in [before[o:o + p.numel()] for p, o in
                                                                       zip(tr.flat.params, tr.flat.offsets)]]))[sel].abs().max())
        err = float((got - want)[sel].abs().max())
        assert step_size > 0 and err < 2e-2 * step_size, (kind, step, err, step_size)


def test_sgd_graph_replay_follows_set_lr():
    """The advisor's finding on r01: the captured SGD step froze lr into the graph.  Now lr / momentum / weight decay come
    from device memory: with lr = 0 a replay must leave the weights alone, with lr > 0 it must move them."""
    from cervix_b200.engine import SegTrainer
    model = _small_model(torch.bfloat16)
    tr = SegTrainer(model, lr=1e-2, optimizer="sgd", momentum=0.9, weight_decay=1e-4, cls_weights=[1, 1, 5, 3, 4])
    imgs, pngs, _ = _batches(1)[0]
    imgs, pngs = imgs.cuda(), pngs.cuda()
    tr.capture(imgs, pngs, None, warmup=1)
    w0 = tr.flat.data.clone()
    tr.set_lr(0.0)
    tr.step_graphed(imgs, pngs)
    assert torch.equal(tr.flat.data, w0)
    tr.set_lr(1e-2)
    tr.step_graphed(imgs, pngs)
    assert float((tr.flat.data - w0).abs().max()) > 1e-5


def test_freeze_then_unfreeze_keeps_optimizer_state_and_skips_frozen_parameters():
    """Freeze_Train (train.py:447-449, 531-551): backbone parameters with requires_grad = False are not touched (not even
    by weight decay), the head trains; after unfreezing the backbone trains too, the head's Adam moments and step count
    carry on, and the backbone starts at Adam step 1 like a parameter torch.optim sees for the first time."""
    from cervix_b200.engine import SegTrainer
    model = _small_model(torch.float32)
    for p in model.backbone.parameters():
        p.requires_grad = False
    tr = SegTrainer(model, lr=1e-3, weight_decay=1e-2, cls_weights=[1, 1, 5, 3, 4])
    assert len(tr.flat.params) == len(list(model.parameters()))       # frozen parameters live in the flat buffer too
    bb = {id(p) for p in model.backbone.parameters()}
    is_bb = torch.zeros(tr.flat.numel, dtype=torch.bool, device="cuda")
    for p, o in zip(tr.flat.params, tr.flat.offsets):
        if id(p) in bb:
            is_bb[o:o + p.numel()] = True
    imgs, pngs, _ = _batches(1)[0]
    imgs, pngs = imgs.cuda(), pngs.cuda()
    w0 = tr.flat.data.clone()
    for _ in range(3):
        tr.step(imgs, pngs)
    moved = (tr.flat.data != w0)
    assert not bool((moved & is_bb).any()) and bool((moved & ~is_bb).any())
    assert float(tr.m[is_bb].abs().max()) == 0.0
    m_head = tr.m[~is_bb].clone()
    for p in model.backbone.parameters():
        p.requires_grad = True
    w1 = tr.flat.data.clone()
    tr.step(imgs, pngs)
    assert bool(((tr.flat.data != w1) & is_bb).any())
    assert len(tr.runs) >= 2 and sorted(int(c) for c in tr.step_devs.tolist()) == [1, 4]
    # the head's first moment moved on from where it was (b1 * m + (1 - b1) * g), it was not reset to (1 - b1) * g
    g_head = tr.flat.grad[~is_bb]
    want_m = 0.9 * m_head + 0.1 * (g_head + 1e-2 * w1[~is_bb])
    assert bool(((tr.m[~is_bb] - want_m).abs() <= 1e-5 * (m_head.abs() + g_head.abs() + 1e-2 * w1[~is_bb].abs()) + 1e-8).all())


def test_trainer_state_roundtrip_resumes_bit_exactly():
    """Missing item 9 of r01: optimizer moments + step counts travel with a checkpoint.  Train 3 steps, save weights +
    trainer state, load both into a fresh model / trainer, take one more step on each: identical parameters."""
    from cervix_b200.engine import SegTrainer
    imgs, pngs, _ = _batches(1)[0]
    imgs, pngs = imgs.cuda(), pngs.cuda()
    # (SGD: its update is linear in the gradient, so the rounding-level run-to-run differences of the weight-gradient
    # atomics stay rounding-level; Adam would turn them into +-lr on parameters whose true gradient is zero)
    kw = dict(lr=1e-2, optimizer="sgd", momentum=0.9, weight_decay=1e-4, cls_weights=[1, 1, 5, 3, 4])
    a = SegTrainer(_small_model(torch.float32), **kw)
    for _ in range(3):
        a.step(imgs, pngs)
    weights = {k: v.clone() for k, v in a.model.state_dict().items()}
    state = a.state_dict()
    fresh = _small_model(torch.float32, seed=5)
    fresh.load_state_dict(weights)
    b = SegTrainer(fresh, **kw).load_state_dict(state)
    assert b.t == 3 and torch.equal(a.m, b.m) and int(b.step_dev) == 3
    la, lb = a.step(imgs, pngs), b.step(imgs, pngs)
    assert torch.equal(la, lb)                        # same weights, fixed-order forward reductions: same losses
    # (the weight-gradient kernels accumulate split-K partials with fp32 atomics: equal to rounding, not to the bit)
    assert torch.allclose(a.flat.data, b.flat.data, rtol=0, atol=1e-5)


def test_two_graph_split_backward_equals_the_single_backward():
    """``capture_split`` (graph A: forward + backward down to the end of Xception's entry flow; graph B: the entry flow's
    backward - the data-parallel step all-reduces A's gradients under B) against the ordinary one-graph step on one GPU.
    The gradients are compared from IDENTICAL weights (lr = 0): on this 4 x 64 x 64 batch the train-mode network amplifies
    a one-ulp weight difference into percent-level gradient changes within a step (two one-graph trainers differ as much,
    tools/debug_split2.py), so trajectories cannot be compared.  The optimizer step of both ranges is then checked against
    SGD written out by hand on the split trainer's own gradient."""
    from cervix_b200.engine import SegTrainer
    imgs, pngs, _ = _batches(1, bsz=4, size=64, seed=3)[0]
    imgs, pngs = imgs.cuda(), pngs.cuda()
    kw = dict(lr=0.0, optimizer="sgd", momentum=0.9, weight_decay=0.0, cls_weights=[1, 1, 5, 3, 4])
    one = SegTrainer(_small_model(torch.float32, seed=2, bb="xception"), **kw).capture(imgs, pngs, None, warmup=1)
    two = SegTrainer(_small_model(torch.float32, seed=2, bb="xception"), **kw).capture_split(imgs, pngs, None, warmup=1)
    assert two is not None and two._split
    (lo_a, hi_a), (lo_b, hi_b) = two._split_ranges
    assert lo_b == 0 and hi_b == lo_a and hi_a == two.flat.numel and 0 < hi_b < 0.1 * hi_a      # the entry flow is small
    assert torch.equal(one.flat.data, two.flat.data)
    for step in range(2):
        la, lb = one.step_graphed(imgs, pngs).clone(), two.step_graphed(imgs, pngs).clone()
        assert torch.allclose(la, lb, rtol=1e-6, atol=1e-7), (step, la, lb)
        ga, gb = one.flat.grad, two.flat.grad
        assert float((ga - gb).abs().max()) <= 1e-5 * float(ga.abs().max()), step          # atomics order of the wgrads
        assert float(gb[:hi_b].abs().max()) > 0 and float(gb[lo_a:].abs().max()) > 0
    assert torch.equal(one.flat.data, two.flat.data)                                        # lr = 0 moved nothing
    # one real step: both ranges are updated, with the gradient of this very replay
    two.set_lr(1e-2)
    w0, m0 = two.flat.data.clone(), two.m.clone()
    two.step_graphed(imgs, pngs)
    g = two.flat.grad
    m1 = 0.9 * m0 + g
    # tolerances relative to the OPERANDS (the kernel fuses multiply-adds; where 0.9 m0 and g cancel, the result's own
    # magnitude says nothing about the rounding)
    assert bool(((two.m - m1).abs() <= 1e-6 * (m0.abs() + g.abs()) + 1e-9).all())
    upd = g + 0.9 * m1 if two.nesterov else m1
    assert bool(((two.flat.data - (w0 - 1e-2 * upd)).abs() <= 1e-6 * (w0.abs() + 1e-2 * (g.abs() + m1.abs())) + 1e-9).all())
    assert float((two.flat.data[:hi_b] - w0[:hi_b]).abs().max()) > 0 and float((two.flat.data[lo_a:] - w0[lo_a:]).abs().max()) > 0
    # a model without a cut (MobileNetV2) declines, and the caller falls back to the one-graph step
    assert SegTrainer(_small_model(torch.float32, bb="mobilenet"), **kw).capture_split(imgs, pngs, None, warmup=1) is None
