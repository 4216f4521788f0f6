"""CPU tests: every autograd shell in ops.py (driven by the emulation backend) against stock
torch autograd on NCHW tensors - forward values and all input/parameter gradients."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import cervix_b200.backend as backend
from cervix_b200 import ops
from tests.emu_backend import EmuBackend


@pytest.fixture(autouse=True)
def emu():
    prev = backend.set_backend(EmuBackend())
    yield
    backend.set_backend(prev)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def close(a, b, tol=1e-4):
    return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("k,stride,pad,dil,bias", [(1, 1, 0, 1, False), (3, 1, 1, 1, True), (3, 1, 6, 6, True),
                                                   (1, 2, 0, 1, False), (3, 2, 1, 1, False)])
def test_conv2d(k, stride, pad, dil, bias):
    torch.manual_seed(0)
    x = torch.randn(2, 8, 13, 11, requires_grad=True)
    conv = nn.Conv2d(8, 16, k, stride, pad, dil, bias=bias)
    y_ref = conv(x); gy = torch.randn_like(y_ref); y_ref.backward(gy)
    ref = (x.grad.clone(), conv.weight.grad.clone(), None if not bias else conv.bias.grad.clone())
    x.grad = None; conv.zero_grad()
    xe = nhwc(x.detach()).requires_grad_(True)
    y = ops.conv2d(xe, conv.weight, conv.bias, stride, pad, dil)
    assert close(nchw(y), y_ref)
    y.backward(nhwc(gy))
    assert close(nchw(xe.grad), ref[0]) and close(conv.weight.grad, ref[1])
    if bias:
        assert close(conv.bias.grad, ref[2])


@pytest.mark.parametrize("stride,dil,relu_in", [(1, 1, True), (2, 1, True), (1, 2, False)])
def test_dwconv(stride, dil, relu_in):
    torch.manual_seed(1)
    x = torch.randn(2, 8, 12, 10, requires_grad=True)
    conv = nn.Conv2d(8, 8, 3, stride, dil, dil, groups=8, bias=False)
    y_ref = conv(F.relu(x) if relu_in else x); gy = torch.randn_like(y_ref); y_ref.backward(gy)
    ref = (x.grad.clone(), conv.weight.grad.clone()); conv.zero_grad()
    xe = nhwc(x.detach()).requires_grad_(True)
    y = ops.dwconv3x3(xe, conv.weight, stride, dil, dil, relu_in)
    assert close(nchw(y), y_ref)
    y.backward(nhwc(gy))
    assert close(nchw(xe.grad), ref[0]) and close(conv.weight.grad, ref[1])


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("act,with_res", [(ops.ACT_NONE, False), (ops.ACT_RELU, True), (ops.ACT_RELU6, False),
                                          (ops.ACT_NONE, True)])
def test_batchnorm_act(training, act, with_res):
    torch.manual_seed(2)
    x = (2 * torch.randn(3, 8, 6, 5) + 1).requires_grad_(True)
    r = torch.randn(3, 8, 6, 5, requires_grad=True) if with_res else None
    bn = nn.BatchNorm2d(8, momentum=0.3)
    with torch.no_grad():
        bn.weight.normal_(1, 0.2); bn.bias.normal_(0, 0.2); bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 1.5)
    bn.train(training)
    bn2 = nn.BatchNorm2d(8, momentum=0.3); bn2.load_state_dict(bn.state_dict()); bn2.train(training)
    y_ref = bn(x)
    if with_res:
        y_ref = y_ref + r
    y_ref = F.relu(y_ref) if act == ops.ACT_RELU else (F.relu6(y_ref) if act == ops.ACT_RELU6 else y_ref)
    gy = torch.randn_like(y_ref); y_ref.backward(gy)
    xe = nhwc(x.detach()).requires_grad_(True)
    re = nhwc(r.detach()).requires_grad_(True) if with_res else None
    y = ops.batchnorm_act(xe, bn2, act, re)
    assert close(nchw(y), y_ref)
    y.backward(nhwc(gy))
    assert close(nchw(xe.grad), x.grad) and close(bn2.weight.grad, bn.weight.grad) and close(bn2.bias.grad, bn.bias.grad)
    if with_res:
        assert close(nchw(re.grad), r.grad)
    assert close(bn2.running_mean, bn.running_mean) and close(bn2.running_var, bn.running_var)
    assert int(bn2.num_batches_tracked) == int(bn.num_batches_tracked)


def test_misc_ops():
    torch.manual_seed(3)
    x = torch.randn(2, 8, 5, 7, requires_grad=True)
    # upsample + to-NCHW upsample
    y_ref = F.interpolate(x, size=(20, 28), mode="bilinear", align_corners=True); gy = torch.randn_like(y_ref)
    y_ref.backward(gy)
    xe = nhwc(x.detach()).requires_grad_(True)
    y = ops.upsample_bilinear(xe, 20, 28); y.backward(nhwc(gy))
    assert close(nchw(y), y_ref) and close(nchw(xe.grad), x.grad)
    xe2 = nhwc(x.detach()).requires_grad_(True)
    y2 = ops.upsample_to_nchw(xe2, 20, 28); y2.backward(gy)
    assert close(y2, y_ref) and close(nchw(xe2.grad), x.grad)
    # global pool -> broadcast
    x.grad = None
    z_ref = x.mean(dim=(2, 3), keepdim=True).expand(2, 8, 5, 7) * 1.0; gz = torch.randn(2, 8, 5, 7); z_ref.backward(gz)
    xe3 = nhwc(x.detach()).requires_grad_(True)
    z = ops.broadcast_hw(ops.global_avg_pool(xe3), 5, 7); z.backward(nhwc(gz))
    assert close(nchw(z), z_ref) and close(nchw(xe3.grad), x.grad)
    # concat + relu
    a = torch.randn(2, 3, 4, 8, requires_grad=True); b = torch.randn(2, 3, 4, 16, requires_grad=True)
    c = ops.relu(ops.cat_channels([a, b])); gc = torch.randn_like(c); c.backward(gc)
    cr = F.relu(torch.cat([a.detach(), b.detach()], 3))
    assert close(c, cr) and close(a.grad, (gc * (cr > 0))[..., :8]) and close(b.grad, (gc * (cr > 0))[..., 8:])


# ---- gradient chains: several consumers of one tensor, summed without separate add passes -----------------------------
def test_grad_chain_hands_over_the_total_once_and_resets():
    from cervix_b200 import ops
    ch = ops.GradChain(3)
    a, b, c = torch.ones(2), 2 * torch.ones(2), 4 * torch.ones(2)
    assert ch.add(a) is None
    assert ch.take(lambda side: side + b) is None            # a consumer that adds the running sum inside its own kernel
    assert torch.equal(ch.add(c), 7 * torch.ones(2))
    assert ch.acc is None and ch.left == 3                   # ready for a second backward over a retained graph
    assert ch.take(lambda side: a if side is None else side + a) is None


def test_aspp_with_and_without_the_gradient_chain(monkeypatch):
    """ASPP's five consumers of the backbone output (deeplabv3_plus.py:89-114): with the chain every consumer but the
    last returns None to autograd and the total arrives once; same input gradient and parameter gradients as autograd's own
    accumulation."""
    import cervix_b200.nets.deeplabv3_plus as dl
    prev = backend.set_backend(EmuBackend())
    try:
        res = {}
        for on in (True, False):
            monkeypatch.setattr(dl, "_GRAD_CHAIN", on)
            torch.manual_seed(0)
            aspp = dl.ASPP(16, 8, rate=1).train()
            x = torch.randn(2, 9, 9, 16).requires_grad_(True)
            y = aspp(x)
            y.mul(torch.linspace(-1, 1, y.numel()).view_as(y)).sum().backward()
            res[on] = (y.detach(), x.grad.clone(), {k: p.grad.clone() for k, p in aspp.named_parameters()})
        assert torch.equal(res[True][0], res[False][0])
        assert torch.allclose(res[True][1], res[False][1], rtol=1e-5, atol=1e-6)
        for k in res[True][2]:
            assert torch.allclose(res[True][2][k], res[False][2][k], rtol=1e-5, atol=1e-6), k
    finally:
        backend.set_backend(prev)
