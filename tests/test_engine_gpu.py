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
