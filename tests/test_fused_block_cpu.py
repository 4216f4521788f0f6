"""CPU tests of the fused Xception block's host logic (cervix_b200/ops_fused.py) over the plain-torch emulation of
the C ABI: the hand-scheduled chain (bn1 folded into the pointwise weights, bn2 deferred to the consumer, bn1's
backward read off the weight-gradient GEMM) must reproduce the operator-by-operator path of ops.py - outputs, input
gradient, every parameter gradient and the BatchNorm running buffers."""
import copy
import os

import pytest
import torch

import cervix_b200.backend as backend
from cervix_b200 import ops_fused
from cervix_b200.nets.xception import Block
from tests.emu_backend import EmuBackend


@pytest.fixture(autouse=True)
def emu():
    prev = backend.set_backend(EmuBackend())
    yield
    backend.set_backend(prev)


def make_block(c, seed, cout=None, stride=1, grow_first=True):
    torch.manual_seed(seed)
    blk = Block(c, cout or c, stride, grow_first=grow_first)
    for m in blk.modules():
        if isinstance(m, torch.nn.Conv2d):
            m.weight.data.normal_(0, (2.0 / (m.weight.shape[1] * 9 if m.groups == 1 else 9)) ** 0.5)
        elif isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.normal_(1, 0.2); m.bias.data.normal_(0, 0.2)
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
    for m in blk.modules():
        m._cervix_dtype = torch.float32
    return blk.train()


def run(blk, x, dy, relu_out, fused, inp_is_relu=True, use_hook=False):
    os.environ["CERVIX_NO_FUSED_BLOCK"] = "0" if fused else "1"
    try:
        xi = x.clone().requires_grad_(True)
        out = blk(xi, inp_is_relu=inp_is_relu, relu_out=relu_out)
        hook = blk.hook_layer
        if use_hook:              # block2's low-level feature takes part in the graph (decoder shortcut)
            (out.mul(dy).sum() + hook.mul(hook.detach().cos()).sum()).backward()
        else:
            out.backward(dy)
    finally:
        os.environ.pop("CERVIX_NO_FUSED_BLOCK", None)
    grads = {k: p.grad.clone() for k, p in blk.named_parameters()}
    bufs = {k: b.clone() for k, b in blk.named_buffers()}
    return out.detach(), xi.grad, grads, bufs


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("relu_out", [True, False])
@pytest.mark.parametrize("shape", [(2, 8, 8, 16), (3, 5, 7, 24)])
def test_fused_identity_block_matches_operator_path(relu_out, shape):
    n, h, w, c = shape
    blk_a = make_block(c, 0)
    blk_b = copy.deepcopy(blk_a)
    g = torch.Generator().manual_seed(1)
    x = torch.relu(torch.randn(shape, generator=g))
    dy = torch.randn(shape, generator=g)
    assert ops_fused.identity_block_fusable(blk_a, x)
    out_a, dx_a, gr_a, bf_a = run(blk_a, x, dy, relu_out, fused=False)
    out_b, dx_b, gr_b, bf_b = run(blk_b, x, dy, relu_out, fused=True)
    assert rel(out_b, out_a) < 1e-5
    assert rel(dx_b, dx_a) < 2e-4
    for k in gr_a:
        if k.endswith("bn1.bias"):
            # d(loss)/d(bn1.bias) is analytically zero (bn2 removes any per-channel shift of the pointwise output);
            # the operator path returns rounding noise, the fused path returns exact zeros
            assert float(gr_b[k].abs().max()) == 0.0
            assert float(gr_a[k].abs().max()) < 1e-3 * float(gr_a[k.replace("bias", "weight")].abs().max() + 1e-6) + 1e-4
            continue
        assert rel(gr_b[k], gr_a[k]) < 5e-4, k
    for k in bf_a:
        assert rel(bf_b[k].float(), bf_a[k].float()) < 1e-5, k


@pytest.mark.parametrize("kind", ["entry_stride2", "exit_skipconv"])
def test_fused_skip_conv_blocks_match_operator_path(kind):
    """Blocks with a 1x1 skip conv: entry flow (blocks 1-3: two fused separable convs, pre-ReLU hook, strided third one
    on the operator path) and block20 (stride 1, grow_first=False: all three fused, skip branch as the residual)."""
    if kind == "entry_stride2":
        blk_a, shape, oshape = make_block(16, 3, 24, 2), (2, 8, 10, 16), (2, 4, 5, 24)
    else:
        blk_a, shape, oshape = make_block(16, 3, 24, 1, grow_first=False), (2, 8, 10, 16), (2, 8, 10, 24)
    blk_b = copy.deepcopy(blk_a)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=g)            # NOT pre-activated: the block applies relu0 itself
    dy = torch.randn(oshape, generator=g)
    out_a, dx_a, gr_a, bf_a = run(blk_a, x, dy, False, fused=False, inp_is_relu=False, use_hook=kind == "entry_stride2")
    out_b, dx_b, gr_b, bf_b = run(blk_b, x, dy, False, fused=True, inp_is_relu=False, use_hook=kind == "entry_stride2")
    assert (blk_b.hook_layer is not None) == (kind == "entry_stride2")
    assert rel(out_b, out_a) < 1e-5 and rel(dx_b, dx_a) < 2e-4
    fused_seps = ("sepconv1", "sepconv2") if kind == "entry_stride2" else ("sepconv1", "sepconv2", "sepconv3")
    for k in gr_a:
        if k.endswith("bn1.bias"):      # analytically zero everywhere; exact zeros on the fused path, noise elsewhere
            if k.split(".")[0] in fused_seps:
                assert float(gr_b[k].abs().max()) == 0.0
            continue
        assert rel(gr_b[k], gr_a[k]) < 5e-4, k
    for k in bf_a:
        assert rel(bf_b[k].float(), bf_a[k].float()) < 1e-5, k


def test_eval_mode_and_no_grad_fall_back():
    blk = make_block(16, 0)
    x = torch.relu(torch.randn(2, 6, 6, 16))
    with torch.no_grad():
        assert not ops_fused.identity_block_fusable(blk, x)
    blk.eval()
    assert not ops_fused.identity_block_fusable(blk, x)
