"""GPU parity tests of the fused separable-conv chain kernels (csrc/dwconv_fused.cu, sepconv.cu, conv_tc.cu's
epilogue extras) against their plain-torch specification (tests/emu_backend.py), and of the whole fused block
against the operator-by-operator CUDA path on the shapes of the Xception middle flow (728 ch, ragged 64-ch chunk)."""
import copy
import os

import pytest
import torch

from cervix_b200 import ops_fused
from cervix_b200.backend import ConvGeom, get_backend
from cervix_b200.nets.xception import Block
from tests.emu_backend import EmuBackend

pytestmark = pytest.mark.gpu
EMU = EmuBackend()


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))


SHAPES = [(2, 32, 32, 728), (3, 17, 21, 64), (2, 8, 40, 136), (1, 64, 64, 256)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("affine", [False, True])
def test_dwf_fwd(shape, affine):
    B = get_backend()
    n, h, w, c = shape
    g = ConvGeom(n, h, w, c, c, 3, 3, 1, 1, 1)
    x = rnd(*shape, seed=1, dtype=torch.bfloat16)
    w9c = rnd(9, c, seed=2, scale=0.3)
    sc = (rnd(c, seed=3, scale=0.3) + 1.0) if affine else None
    sh = rnd(c, seed=4, scale=0.5) if affine else None
    y, st = B.dwf_fwd(x, w9c, sc, sh, True, g, True)
    yr, str_ = EMU.dwf_fwd(x, w9c, sc, sh, True, g, True)
    assert rel(y.float(), yr.float()) < 1e-2
    # statistics are taken over the kernel's own (bf16-rounded) output
    own = EMU._stats(y)
    assert rel(st, own) < 1e-5
    assert rel(st, str_) < 2e-2


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("affine,with_addend,side", [(False, True, True), (True, False, True), (True, False, False),
                                                     (False, False, False)])
def test_dwf_bwd(shape, affine, with_addend, side):
    B = get_backend()
    n, h, w, c = shape
    g = ConvGeom(n, h, w, c, c, 3, 3, 1, 1, 1)
    x = rnd(*shape, seed=1, dtype=torch.bfloat16)
    dd = rnd(*shape, seed=5, dtype=torch.bfloat16)
    w9c = rnd(9, c, seed=2, scale=0.3)
    sc = (rnd(c, seed=3, scale=0.3) + 1.0) if affine else None
    sh = rnd(c, seed=4, scale=0.5) if affine else None
    add = rnd(*shape, seed=6, dtype=torch.bfloat16) if with_addend else None
    dside = rnd(*shape, seed=7, dtype=torch.bfloat16) if side else None
    negk = rnd(c, seed=8, scale=0.3) if side else None
    kmean = rnd(c, seed=9, scale=0.3) if side else None
    gx, dw, sums = B.dwf_bwd(dd, dside, negk, kmean, x, w9c, sc, sh, True, add, g, True)
    gxr, dwr, sumsr = EMU.dwf_bwd(dd, dside, negk, kmean, x, w9c, sc, sh, True, add, g, True)
    assert rel(gx.float(), gxr.float()) < 1e-2
    assert rel(dw, dwr) < 2e-3
    assert rel(sums, sumsr) < 2e-2
    if not with_addend:       # the reduction is over the kernel's own stored gradient
        a = gx.float().reshape(-1, c).double(); b = x.float().reshape(-1, c).double()
        assert rel(sums, torch.stack([a.sum(0), (a * b).sum(0)])) < 1e-5


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 32, 32, 728, 728), (1, 64, 64, 256, 728), (2, 24, 24, 128, 256),
                                            (1, 32, 32, 728, 1024), (3, 16, 16, 64, 128)])
def test_conv_epilogue_side_and_stats(n, h, w, cin, cout):
    B = get_backend()
    g = ConvGeom(n, h, w, cin, cout, 1, 1, 1, 0, 1)
    x = rnd(n, h, w, cin, seed=1, dtype=torch.bfloat16)
    wt = rnd(cout, cin, 1, 1, seed=2, scale=(1.0 / cin) ** 0.5)
    bias = rnd(cout, seed=3)
    wp = B.pack_weight(wt, torch.bfloat16, False)
    y, st = B.conv_fwd_ex(x, wp, bias, g, None, None, True)
    yr, _ = EMU.conv_fwd_ex(x, wp, bias, g, None, None, False)
    assert rel(y.float(), yr.float()) < 1.2e-2
    assert rel(st, EMU._stats(y)) < 1e-5
    # forward with side input AND statistics (both epilogue extras at once)
    sidef = rnd(n, h, w, cout, seed=8, dtype=torch.bfloat16)
    ssf = rnd(cout, seed=9)
    y2, st2 = B.conv_fwd_ex(x, wp, bias, g, sidef, ssf, True)
    y2r, _ = EMU.conv_fwd_ex(x, wp, bias, g, sidef, ssf, False)
    assert rel(y2.float(), y2r.float()) < 1.2e-2
    assert rel(st2, EMU._stats(y2)) < 1e-5
    # data gradient with the side term: dx = dy . W^T + bias + side_scale * side
    dy = rnd(n, h, w, cout, seed=4, dtype=torch.bfloat16)
    side = rnd(n, h, w, cin, seed=5, dtype=torch.bfloat16)
    ss, bi = rnd(cin, seed=6), rnd(cin, seed=7)
    wpt = B.pack_weight(wt, torch.bfloat16, True)
    dx = B.conv_dgrad_ex(dy, wpt, g, bi, side, ss)
    dxr = EMU.conv_dgrad_ex(dy, wpt, g, bi, side, ss)
    assert rel(dx.float(), dxr.float()) < 1.2e-2


@pytest.mark.parametrize("n,h,w,cin,cout,k,stride", [(2, 16, 16, 256, 1024, 1, 1), (3, 20, 12, 64, 64, 3, 1),
                                                      (2, 32, 32, 128, 128, 3, 2), (1, 8, 8, 1024, 256, 1, 1)])
def test_conv_act_epilogue(n, h, w, cin, cout, k, stride):
    """cvx_conv_fwd_tc_act: act(conv + bias + side_scale*side) - the folded conv/bn/residual/relu group of the
    classifier's ResNet-101 encoder; the stride-2 3x3 takes the single-CTA kernel, whose epilogue has the ReLU too."""
    B = get_backend()
    g = ConvGeom(n, h, w, cin, cout, k, k, stride, k // 2, 1)
    x = rnd(n, h, w, cin, seed=1, dtype=torch.bfloat16)
    wt = rnd(cout, cin, k, k, seed=2, scale=(1.0 / (cin * k * k)) ** 0.5)
    bias = rnd(cout, seed=3)
    wp = B.pack_weight(wt, torch.bfloat16, False)
    y = B.conv_fwd_act(x, wp, bias, g, 1)
    yr = EMU.conv_fwd_act(x, wp, bias, g, 1)
    assert float(y.float().min()) >= 0.0 and rel(y.float(), yr.float()) < 1.2e-2
    if stride == 1:
        side = rnd(n, g.ho, g.wo, cout, seed=4, dtype=torch.bfloat16)
        ss = rnd(cout, seed=5).abs() + 0.5
        y2 = B.conv_fwd_act(x, wp, bias, g, 1, side, ss)
        y2r = EMU.conv_fwd_act(x, wp, bias, g, 1, side, ss)
        assert rel(y2.float(), y2r.float()) < 1.2e-2
        y4 = B.conv_fwd_act(x, wp, bias, g, 1, side, None)          # plain residual add
        assert rel(y4.float(), EMU.conv_fwd_act(x, wp, bias, g, 1, side, None).float()) < 1.2e-2
        y3 = B.conv_fwd_act(x, wp, None, g, 0, side, ss)
        assert rel(y3.float(), EMU.conv_fwd_act(x, wp, None, g, 0, side, ss).float()) < 1.2e-2


def test_small_kernels():
    B = get_backend()
    c, cout, rows = 728, 1024, 4096
    x = rnd(rows, c, seed=1, dtype=torch.bfloat16).reshape(4, 32, 32, c)
    st = B.bn_stats(x)
    assert rel(st, EMU.bn_stats(x)) < 1e-6
    gam, bet = rnd(c, seed=2) + 1, rnd(c, seed=3)
    rm, rv = rnd(c, seed=4), rnd(c, seed=5).abs() + 0.5
    rm2, rv2 = rm.clone(), rv.clone()
    off = rnd(c, seed=11)
    got = B.bn_affine(st, rows, gam, bet, rm, rv, 0.1, 1e-5, off)
    ref = EMU.bn_affine(st, rows, gam, bet, rm2, rv2, 0.1, 1e-5, off)
    for a, b in zip(got, ref):
        assert rel(a, b) < 1e-5
    assert rel(rm, rm2) < 1e-6 and rel(rv, rv2) < 1e-6
    wt = rnd(cout, c, 1, 1, seed=6)
    wp, wpt, bias = B.pw_fold(wt, got[2], got[3], torch.bfloat16)
    wpr, wptr, biasr = EMU.pw_fold(wt, got[2], got[3], torch.bfloat16)
    assert torch.equal(wp, wpr) and torch.equal(wpt, wptr) and rel(bias, biasr) < 1e-5
    res = rnd(4, 32, 32, c, seed=7, dtype=torch.bfloat16)
    for act in (0, 1):
        y = B.affine_act(x, got[2], got[3], res, act)
        assert rel(y.float(), EMU.affine_act(x, got[2], got[3], res, act).float()) < 1e-2
    dy = rnd(4, 32, 32, c, seed=8, dtype=torch.bfloat16)
    yact = B.affine_act(x, got[2], got[3], res, 1)
    for act, yy in ((0, None), (1, yact)):
        sums = B.bn_bwd_sums(dy, yy, x, act)
        assert rel(sums, EMU.bn_bwd_sums(dy, yy, x, act)) < 1e-6
        co = B.bn_bwd_coef(sums, rows, got[0], got[1], gam)
        cor = EMU.bn_bwd_coef(sums, rows, got[0], got[1], gam)
        for a, b in zip(co, cor):
            assert rel(a, b) < 1e-4
        dp, gg = B.bn_bwd_affine(dy, yy, x, co[0], co[1], co[2], act, True)
        dpr, ggr = EMU.bn_bwd_affine(dy, yy, x, co[0], co[1], co[2], act, True)
        assert rel(dp.float(), dpr.float()) < 1e-2 and torch.equal(gg, ggr)
    G = rnd(1, cout, c, seed=9)
    got2 = B.pw_bwd_coef(G, wt, got[2], got[1], got[0], rows)
    ref2 = EMU.pw_bwd_coef(G, wt, got[2], got[1], got[0], rows)
    for a, b in zip(got2, ref2):
        if float(b.abs().max()) == 0:
            assert float(a.abs().max()) == 0
        else:
            assert rel(a, b) < 1e-4


def make_block(c, seed, dtype):
    torch.manual_seed(seed)
    blk = Block(c, c, 1)
    for m in blk.modules():
        if isinstance(m, torch.nn.Conv2d):
            m.weight.data.normal_(0, (2.0 / (m.weight.shape[1] * 9 if m.groups == 1 else 9)) ** 0.5)
        elif isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.normal_(1, 0.2); m.bias.data.normal_(0, 0.2)
    return blk.cuda().train()


def run(blk, x, dy, relu_out, fused):
    os.environ["CERVIX_NO_FUSED_BLOCK"] = "0" if fused else "1"
    try:
        xi = x.clone().requires_grad_(True)
        out = blk(xi, inp_is_relu=True, relu_out=relu_out)
        out.backward(dy)
    finally:
        os.environ.pop("CERVIX_NO_FUSED_BLOCK", None)
    res = {"out": out.detach().float(), "dx": xi.grad.float()}
    res.update({"grad:" + k: p.grad.clone() for k, p in blk.named_parameters()})
    res.update({"buf:" + k: b.clone().float() for k, b in blk.named_buffers() if "num_batches" not in k})
    return res


def err(a, ref):
    return float((a.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-20))


@pytest.mark.parametrize("relu_out", [True, False])
@pytest.mark.parametrize("shape", [(4, 32, 32, 728), (2, 16, 16, 128), (8, 32, 32, 256)])
def test_fused_block_is_as_accurate_as_operator_path_bf16(relu_out, shape):
    """Three runs of the same block on the same data: the fp32 operator path (SIMT kernels) as the yardstick, the
    bf16 operator path and the bf16 fused path.  The two bf16 paths round at different points, so they are not
    compared with each other: the fused path must be as close to fp32 as the operator path is (within 1.5x + a
    floor), tensor by tensor - output, input gradient, every parameter gradient, every running buffer."""
    n, h, w, c = shape
    blk_ref = make_block(c, 0, torch.float32)
    blk_a, blk_b = copy.deepcopy(blk_ref), copy.deepcopy(blk_ref)
    x32 = torch.relu(rnd(*shape, seed=1)).bfloat16().float()
    dy32 = rnd(*shape, seed=2).bfloat16().float()
    ref = run(blk_ref, x32, dy32, relu_out, fused=False)
    assert ops_fused.identity_block_fusable(blk_b, x32.bfloat16())
    a = run(blk_a, x32.bfloat16(), dy32.bfloat16(), relu_out, fused=False)
    b = run(blk_b, x32.bfloat16(), dy32.bfloat16(), relu_out, fused=True)
    for k, r in ref.items():
        if k.endswith("bn1.bias") and k.startswith("grad:"):
            assert float(b[k].abs().max()) == 0.0      # analytically zero (see sepconv.cu)
            continue
        ea, eb = err(a[k].float(), r), err(b[k].float(), r)
        assert eb < 1.5 * ea + 5e-3, (k, ea, eb)


@pytest.mark.parametrize("case", [
    # n, h, w, cin, cout, k, stride, pad, dil, act, residual, bias
    (2, 24, 28, 304, 256, 3, 1, 1, 1, 1, False, True),     # decoder cat_conv
    (4, 16, 16, 2048, 256, 3, 1, 6, 6, 1, False, True),    # ASPP atrous branch
    (2, 32, 32, 256, 48, 1, 1, 0, 1, 1, False, True),      # decoder shortcut
    (2, 32, 32, 128, 256, 1, 2, 0, 1, 0, False, False),    # strided 1x1 skip conv + skipbn
    (4, 16, 16, 728, 1024, 1, 1, 0, 1, 0, True, False),    # pointwise + bn2 + residual (exit flow)
    (2, 40, 40, 32, 64, 3, 1, 1, 1, 1, False, False),      # stem conv2
])
def test_conv_bn_act_composite_matches_separate_ops(case):
    """ops.conv_bn_act (BatchNorm statistics from the GEMM epilogue) against ops.conv2d + ops.batchnorm_act on the same
    bf16 inputs: output, running buffers, and all gradients."""
    from cervix_b200 import ops
    n, h, w, cin, cout, k, stride, pad, dil, act, with_res, with_bias = case
    x0 = rnd(n, h, w, cin, seed=1, dtype=torch.bfloat16)
    wt0 = rnd(cout, cin, k, k, seed=2, scale=(cin * k * k) ** -0.5)
    b0 = rnd(cout, seed=3, scale=0.3) if with_bias else None
    ho = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
    wo = (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    r0 = rnd(n, ho, wo, cout, seed=4, dtype=torch.bfloat16) if with_res else None
    dy = rnd(n, ho, wo, cout, seed=5, dtype=torch.bfloat16)
    outs = []
    for fused in (True, False):
        ops._CONV_BN_MIN_K = 0 if fused else (1 << 30)          # (the composite is off by default: see ops.conv_bn_act)
        bn = torch.nn.BatchNorm2d(cout, momentum=0.1).cuda().train()
        with torch.no_grad():
            bn.weight.copy_(1 + 0.2 * rnd(cout, seed=6)); bn.bias.copy_(0.2 * rnd(cout, seed=7))
        x = x0.clone().requires_grad_(True)
        wt = wt0.clone().requires_grad_(True)
        b = None if b0 is None else b0.clone().requires_grad_(True)
        r = None if r0 is None else r0.clone().requires_grad_(True)
        if fused:
            y = ops.conv_bn_act(x, wt, stride, pad, dil, bn, act, r, b)
            assert type(y.grad_fn).__name__.startswith("ConvBnAct")
        else:
            y = ops.conv_bn_act(x, wt, stride, pad, dil, bn, act, r, b)
            assert type(y.grad_fn).__name__.startswith("BatchNormAct")
        ops._CONV_BN_MIN_K = 1 << 30
        y.backward(dy)
        outs.append(dict(y=y.detach(), dx=x.grad, dw=wt.grad, dg=bn.weight.grad, db=bn.bias.grad, rm=bn.running_mean.clone(),
                         rv=bn.running_var.clone(), dr=None if r is None else r.grad, nbt=int(bn.num_batches_tracked)))
    a, b_ = outs
    assert a["nbt"] == b_["nbt"] == 1
    assert rel(a["rm"], b_["rm"]) < 1e-4 and rel(a["rv"], b_["rv"]) < 1e-4
    assert rel(a["y"].float(), b_["y"].float()) < 1e-2          # one bf16 ulp where scale/shift differ in the last bit
    assert rel(a["dx"].float(), b_["dx"].float()) < 2e-2
    assert rel(a["dw"], b_["dw"]) < 2e-2
    assert rel(a["dg"], b_["dg"]) < 1e-2 and rel(a["db"], b_["db"]) < 1e-2
    if with_res:
        assert rel(a["dr"].float(), b_["dr"].float()) < 1e-2
