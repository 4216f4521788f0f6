"""GPU parity tests, kernel by kernel: every C-ABI entry point (through backend.CudaBackend)
against its plain-torch specification (tests/emu_backend.py) on the same seeded inputs.

Tolerances: fp32 storage -> 2e-4 of the reference's max magnitude (fp32 accumulation order);
bf16 storage -> 1.2e-2 (one bf16 rounding of the output, inputs are identical bf16 values).
"""
import pytest
import torch

from cervix_b200 import _lib
from cervix_b200.backend import ConvGeom, get_backend
from tests.emu_backend import EmuBackend

pytestmark = pytest.mark.gpu

EMU = EmuBackend()
DTYPES = [torch.float32, torch.bfloat16]


@pytest.fixture(scope="module")
def B():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return get_backend()


def tol(dtype):
    return 2e-4 if dtype == torch.float32 else 1.2e-2


def check(a, b, t, what=""):
    a = a.float(); b = b.float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = float((a - b).abs().max())
    ref = max(float(b.abs().max()), 1e-6)
    assert err <= t * ref, "%s: max err %.4e vs ref max %.4e (tol %.1e)" % (what, err, ref, t)


def rnd(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


def test_loads_native_library(B):
    assert B.name == "cuda" and B.lib.cvx_abi_version() == 1
    assert B.is_sm100(), "these tests are meant for a B200 (sm_100)"


@pytest.mark.parametrize("dtype", DTYPES)
def test_layout_roundtrip(B, dtype):
    x = rnd((3, 5, 17, 33), torch.float32, 0)
    y = B.to_nhwc(x, dtype)
    check(y, EMU.to_nhwc(x, dtype), 1e-6 if dtype == torch.float32 else 4e-3)
    check(B.to_nchw(y), EMU.to_nchw(y), 1e-6)
    w = rnd((24, 16, 3, 3), torch.float32, 1)
    for tf in (False, True):
        check(B.pack_weight(w, dtype, tf), EMU.pack_weight(w, dtype, tf), 1e-6 if dtype == torch.float32 else 4e-3)
    gp = rnd((9, 24, 16), torch.float32, 2)
    check(B.unpack_wgrad(gp, 24, 16, 3, 3), EMU.unpack_wgrad(gp, 24, 16, 3, 3), 0)
    wd = rnd((40, 1, 3, 3), torch.float32, 3)
    check(B.pack_dw_weight(wd), EMU.pack_dw_weight(wd), 0)
    check(B.unpack_dw_wgrad(B.pack_dw_weight(wd)), wd, 0)
    xs = [rnd((2, 7, 9, c), dtype, 10 + i) for i, c in enumerate((8, 24, 5))]
    cat = B.cat_channels(xs)
    check(cat, EMU.cat_channels(xs), 0)
    check(B.slice_channels(cat, 8, 24), xs[1], 0)
    check(B.slice_channels(cat, 32, 5), xs[2], 0)


CONV_CASES = [
    # n, h, w, cin, cout, k, stride, pad, dil, bias
    (2, 19, 23, 3, 32, 3, 2, 1, 1, False),     # stem
    (2, 33, 64, 3, 32, 3, 2, 1, 1, False),     # stem, output rows of 32 pixels (whole 16-pixel MMA blocks per row)
    (1, 19, 21, 3, 32, 3, 2, 1, 1, False),     # stem, 110 output pixels (ragged last MMA block)
    (2, 16, 16, 32, 64, 3, 1, 1, 1, False),
    (2, 16, 12, 64, 128, 1, 2, 0, 1, False),   # strided skip
    (1, 12, 12, 72, 40, 3, 1, 6, 6, True),     # atrous, ragged channels
    (3, 9, 7, 40, 5, 1, 1, 0, 1, True),        # classifier
    (2, 13, 11, 256, 5, 1, 1, 0, 1, True),     # classifier at the real channel count (register-resident fast paths)
    (1, 5, 3, 128, 5, 1, 1, 0, 1, False),
    (1, 8, 8, 24, 24, 3, 2, 1, 1, False),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_simt(B, dtype, case):
    n, h, w, cin, cout, k, s, p, d, bias = case
    g = ConvGeom(n, h, w, cin, cout, k, k, s, p, d)
    x = rnd((n, h, w, cin), dtype, 1)
    wt = rnd((cout, cin, k, k), torch.float32, 2, 0.2)
    b = rnd((cout,), torch.float32, 3) if bias else None
    wp, wpt = B.pack_weight(wt, dtype, False), B.pack_weight(wt, dtype, True)
    check(B.conv_fwd(x, wp, b, g, False), EMU.conv_fwd(x, wp, b, g, False), tol(dtype), "fwd")
    dy = rnd((n, g.ho, g.wo, cout), dtype, 4)
    check(B.conv_dgrad(dy, wpt, g, False), EMU.conv_dgrad(dy, wpt, g, False), tol(dtype), "dgrad")
    check(B.conv_wgrad(x, dy, g, False), EMU.conv_wgrad(x, dy, g, False), 5e-4 if dtype == torch.float32 else 2e-3, "wgrad")
    check(B.bias_grad(dy), EMU.bias_grad(dy), 1e-4, "bias_grad")


@pytest.mark.parametrize("dtype", DTYPES)
def test_subsample(B, dtype):
    x = rnd((2, 9, 12, 16), dtype, 5)
    y = B.subsample(x, 2)
    check(y, EMU.subsample(x, 2), 0)
    check(B.subsample_bwd(y, 9, 12, 2), EMU.subsample_bwd(y, 9, 12, 2), 0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("c,stride,dil,relu_in", [(64, 1, 1, True), (728, 1, 1, True), (128, 2, 1, True),
                                                  (144, 1, 2, False), (96, 2, 1, False)])
def test_depthwise(B, dtype, c, stride, dil, relu_in):
    n, h, w = 2, 13, 18
    g = ConvGeom(n, h, w, c, c, 3, 3, stride, dil, dil)
    x = rnd((n, h, w, c), dtype, 1)
    w9c = B.pack_dw_weight(rnd((c, 1, 3, 3), torch.float32, 2, 0.3))
    check(B.dw_fwd(x, w9c, g, relu_in), EMU.dw_fwd(x, w9c, g, relu_in), tol(dtype), "fwd")
    dy = rnd((n, g.ho, g.wo, c), dtype, 3)
    check(B.dw_bwd_data(dy, w9c, x, g, relu_in), EMU.dw_bwd_data(dy, w9c, x, g, relu_in), tol(dtype), "bwd_data")
    check(B.dw_bwd_weight(x, dy, g, relu_in), EMU.dw_bwd_weight(x, dy, g, relu_in), 1e-3, "bwd_weight")


@pytest.mark.parametrize("n,h,w,c", [(4, 128, 128, 64), (3, 40, 70, 136), (2, 32, 32, 728)])
def test_depthwise_tiled_pipeline(B, n, h, w, c):
    """bf16 stride-1 path (TMA-staged tiles): many tiles per CTA, ragged image and channel tails."""
    dtype = torch.bfloat16
    g = ConvGeom(n, h, w, c, c, 3, 3, 1, 1, 1)
    x = rnd((n, h, w, c), dtype, 1)
    w9c = B.pack_dw_weight(rnd((c, 1, 3, 3), torch.float32, 2, 0.3))
    dy = rnd((n, h, w, c), dtype, 3)
    for relu_in in (True, False):
        check(B.dw_fwd(x, w9c, g, relu_in), EMU.dw_fwd(x, w9c, g, relu_in), tol(dtype), "fwd")
        check(B.dw_bwd_data(dy, w9c, x, g, relu_in), EMU.dw_bwd_data(dy, w9c, x, g, relu_in), tol(dtype), "bwd_data")
        check(B.dw_bwd_weight(x, dy, g, relu_in), EMU.dw_bwd_weight(x, dy, g, relu_in), 2e-3, "bwd_weight")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("c,act,res", [(64, 0, False), (728, 1, True), (48, 2, False), (2048, 1, False), (304, 0, True)])
def test_batchnorm(B, dtype, training, c, act, res):
    shape = (2, 11, 9, c)
    x = (rnd(shape, torch.float32, 1) * 2 + 0.7).to(dtype)
    r = rnd(shape, dtype, 2) if res else None
    gamma = 1 + 0.2 * rnd((c,), torch.float32, 3); beta = 0.2 * rnd((c,), torch.float32, 4)
    rm = 0.3 * rnd((c,), torch.float32, 5); rv = 0.5 + torch.rand(c, device="cuda")
    rm2, rv2 = rm.clone(), rv.clone()
    y, mean, invstd = B.bn_forward(x, r, gamma, beta, rm, rv, act, training, 0.1, 1e-5)
    ye, me, ie = EMU.bn_forward(x, r, gamma, beta, rm2, rv2, act, training, 0.1, 1e-5)
    check(y, ye, tol(dtype), "y"); check(mean, me, 1e-5, "mean"); check(invstd, ie, 1e-5, "invstd")
    check(rm, rm2, 1e-5, "running_mean"); check(rv, rv2, 1e-5, "running_var")
    dy = rnd(shape, dtype, 6)
    out = B.bn_backward(dy, x, ye, gamma, me, ie, act, training, res)
    oute = EMU.bn_backward(dy, x, ye, gamma, me, ie, act, training, res)
    check(out[0], oute[0], tol(dtype) * 2, "dx")
    if res:
        check(out[1], oute[1], tol(dtype), "dres")
    check(out[2], oute[2], 1e-4, "dgamma"); check(out[3], oute[3], 1e-4, "dbeta")
    if act != 0 and not res:
        # mask recomputed from x (beta given, y not read) against the same recipe in fp32 torch ops ...
        outx = B.bn_backward(dy, x, None, gamma, mean, invstd, act, training, False, beta)
        outs = EMU.bn_backward(dy, x, None, gamma, me, ie, act, training, False, beta)
        check(outx[0], outs[0], tol(dtype) * 2, "dx (mask from x)")
        check(outx[2], outs[2], 2e-4, "dgamma (mask from x)"); check(outx[3], outs[3], 2e-4, "dbeta (mask from x)")
        if act == 1 or dtype == torch.float32:
            # ... and identical to the mask read off the product's own y (ReLU6 in bf16 differs where y rounds to 6.0:
            # the recomputed mask is the fp32 one, as torch's hardtanh_backward uses the pre-activation value)
            outy = B.bn_backward(dy, x, y, gamma, mean, invstd, act, training, False)
            check(outx[0], outy[0], 1e-6, "dx (mask from x vs y)")
            check(outx[2], outy[2], 1e-6, "dgamma (mask from x vs y)"); check(outx[3], outy[3], 1e-6, "dbeta (mask from x vs y)")


def test_batchnorm_tiny_batch_rows(B):
    # ASPP image-pooling branch: BatchNorm over a [B,1,1,256] tensor (statistics over the batch only)
    x = rnd((4, 1, 1, 256), torch.float32, 1)
    gamma = torch.ones(256, device="cuda"); beta = torch.zeros(256, device="cuda")
    rm = torch.zeros(256, device="cuda"); rv = torch.ones(256, device="cuda")
    y, m, i = B.bn_forward(x, None, gamma, beta, rm, rv, 1, True, 0.1, 1e-5)
    ye, me, ie = EMU.bn_forward(x, None, gamma, beta, rm.clone().zero_(), rv.clone().fill_(1), 1, True, 0.1, 1e-5)
    check(y, ye, 1e-4); check(i, ie, 1e-5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_small_ops(B, dtype):
    x = rnd((2, 6, 10, 48), dtype, 1)
    check(B.relu_fwd(x), EMU.relu_fwd(x), 0)
    dy = rnd(x.shape, dtype, 2)
    check(B.relu_bwd(dy, EMU.relu_fwd(x)), EMU.relu_bwd(dy, EMU.relu_fwd(x)), 0)
    check(B.add(x, dy), EMU.add(x, dy), 4e-3)
    check(B.spatial_reduce(x, 1 / 60), EMU.spatial_reduce(x, 1 / 60), tol(dtype))
    g = rnd((2, 1, 1, 48), dtype, 3)
    check(B.spatial_broadcast(g, 6, 10, 0.5), EMU.spatial_broadcast(g, 6, 10, 0.5), 4e-3)
    check(B.upsample_fwd(x, 24, 40), EMU.upsample_fwd(x, 24, 40), tol(dtype), "up fwd")
    check(B.upsample_fwd(x, 21, 37), EMU.upsample_fwd(x, 21, 37), tol(dtype), "up fwd odd")
    dyu = rnd((2, 24, 40, 48), dtype, 4)
    check(B.upsample_bwd(dyu, 6, 10), EMU.upsample_bwd(dyu, 6, 10), tol(dtype), "up bwd")
    dyo = rnd((2, 21, 37, 48), dtype, 5)
    check(B.upsample_bwd(dyo, 6, 10), EMU.upsample_bwd(dyo, 6, 10), tol(dtype), "up bwd odd")
    # the decoder junction: upsample written into / its gradient read from a channel slice of the concat buffer -
    # bit-identical to upsample followed by cat (same arithmetic, different addresses)
    tail = rnd((2, 24, 40, 24), dtype, 8)
    cat = B.upsample_concat(x, tail)
    assert torch.equal(cat, torch.cat([B.upsample_fwd(x, 24, 40), tail], dim=3))
    dcat = rnd((2, 24, 40, 72), dtype, 9)
    dxc, dtail = B.upsample_concat_bwd(dcat, 48, 6, 10)
    assert torch.equal(dxc, B.upsample_bwd(dcat[..., :48].contiguous(), 6, 10)) and torch.equal(dtail, dcat[..., 48:])
    z = rnd((2, 6, 10, 5), dtype, 6)
    check(B.upsample_to_nchw_fwd(z, 24, 40), EMU.upsample_to_nchw_fwd(z, 24, 40), 1e-5 if dtype == torch.float32 else 1e-5, "to_nchw fwd")
    dz = rnd((2, 5, 24, 40), torch.float32, 7)
    check(B.upsample_to_nchw_bwd(dz, 6, 10, dtype), EMU.upsample_to_nchw_bwd(dz, 6, 10, dtype), tol(dtype), "to_nchw bwd")
    # separable row kernel (up-sampling) on odd sizes, rows wider than a block, the identity size; gather kernel otherwise
    for shape, (hi, wi) in (((3, 5, 37, 53), (10, 14)), ((1, 5, 128, 128), (32, 32)), ((1, 3, 40, 300), (10, 150)),
                            ((2, 5, 6, 10), (6, 10)), ((1, 5, 6, 10), (12, 20)), ((1, 2, 64, 64), (1, 1))):
        dz = rnd(shape, torch.float32, 8)
        check(B.upsample_to_nchw_bwd(dz, hi, wi, dtype), EMU.upsample_to_nchw_bwd(dz, hi, wi, dtype), tol(dtype),
              "to_nchw bwd %s -> %dx%d" % (shape, hi, wi))


@pytest.mark.parametrize("dtype", DTYPES)
def test_maxpool_and_im2col(B, dtype):
    x = rnd((2, 13, 18, 64), dtype, 1)
    y = B.maxpool_fwd(x)
    check(y, EMU.maxpool_fwd(x), 0, "maxpool fwd")
    dy = rnd(tuple(y.shape), dtype, 2)
    check(B.maxpool_bwd(x, y, dy), EMU.maxpool_bwd(x, y, dy), 1e-6 if dtype == torch.float32 else 8e-3, "maxpool bwd")
    img = rnd((2, 21, 30, 3), dtype, 3)
    g = ConvGeom(2, 21, 30, 3, 64, 7, 7, 2, 3, 1)
    check(B.im2col_narrow(img, g, 152), EMU.im2col_narrow(img, g, 152), 0, "im2col")


@pytest.mark.parametrize("dtype", DTYPES)
def test_dropout(B, dtype):
    x = rnd((4, 32, 32, 64), dtype, 1) + 3
    y, mask = B.dropout_fwd(x, 0.5, 1234)
    keep = float(mask.float().mean())
    assert abs(keep - 0.5) < 0.01
    check(y, (x.float() * mask * 2).to(dtype), 4e-3)
    y2, mask2 = B.dropout_fwd(x, 0.5, 1234)
    assert torch.equal(mask, mask2)
    _, mask3 = B.dropout_fwd(x, 0.5, 1235)
    assert not torch.equal(mask, mask3)
    dy = rnd(x.shape, dtype, 2)
    check(B.dropout_bwd(dy, mask, 0.5), EMU.dropout_bwd(dy, mask, 0.5), 4e-3)
    # the seeded backward recomputes the keep decisions: identical to the masked one; no mask is written on request
    assert torch.equal(B.dropout_bwd_seeded(dy, 0.5, 1234), B.dropout_bwd(dy, mask, 0.5))
    y_nomask, none = B.dropout_fwd(x, 0.5, 1234, None, want_mask=False)
    assert none is None and torch.equal(y_nomask, y)
    step = torch.tensor([3], dtype=torch.int32, device="cuda")
    ys, ms = B.dropout_fwd(x, 0.5, 1234, step)
    assert not torch.equal(ms, mask) and torch.equal(B.dropout_bwd_seeded(dy, 0.5, 1234, step), B.dropout_bwd(dy, ms, 0.5))
    odd = x.reshape(-1)[:1003].contiguous()          # ragged length: vector body + scalar tail
    yo, mo = B.dropout_fwd(odd, 0.5, 99)
    assert torch.equal(yo, (odd.float() * mo / 0.5).to(odd.dtype)) and 0.4 < float(mo.float().mean()) < 0.6
    _, m1 = B.dropout_fwd(x, 0.1, 7)
    assert abs(float(m1.float().mean()) - 0.9) < 0.01


@pytest.mark.parametrize("with_onehot", [True, False])
def test_seg_loss(B, with_onehot):
    n, c, h, w = 3, 5, 37, 41
    logits = rnd((n, c, h, w), torch.float32, 1, 3.0)
    g = torch.Generator(device="cuda").manual_seed(2)
    target = torch.randint(0, c + 1, (n, h, w), generator=g, device="cuda")
    onehot = torch.eye(c + 1, device="cuda")[target] if with_onehot else None
    cls_w = torch.tensor([1, 1, 5, 3, 4], dtype=torch.float32, device="cuda")
    stats = B.seg_loss_stats(logits, target, onehot, cls_w, 0.5, 2.0, 0.5)
    se = EMU.seg_loss_stats(logits, target, onehot, cls_w, 0.5, 2.0, 0.5)
    assert float(((stats - se).abs() / se.abs().clamp_min(1.0)).max()) < 2e-5
    res = B.seg_loss_finalize(stats, c, 1.0, 1e-5)
    check(res, EMU.seg_loss_finalize(se, c, 1.0, 1e-5), 2e-5, "losses")
    for gv in ([1, 0, 0], [0, 1, 0], [0, 0, 1], [0.3, 1.0, 1.0]):
        gup = torch.tensor(gv, dtype=torch.float32, device="cuda")
        d = B.seg_loss_grad(logits, target, onehot, cls_w, stats, gup, 0.5, 2.0, 1.0, 1e-5)
        de = EMU.seg_loss_grad(logits, target, onehot, cls_w, se, gup, 0.5, 2.0, 1.0, 1e-5)
        check(d, de, 2e-4, "dlogits %s" % gv)


def test_optimizers(B):
    n = 100003
    p = rnd((n,), torch.float32, 1); g = rnd((n,), torch.float32, 2)
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    ref = torch.nn.Parameter(p.clone()); opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), weight_decay=5e-4)
    for t in range(1, 4):
        ref.grad = g.clone() * t
        opt.step()
        B.adam_step(p, g * t, m, v, 1e-3, 0.9, 0.999, 1e-8, 5e-4, t)
    check(p, ref.detach(), 1e-5, "adam")
    p = rnd((n,), torch.float32, 3); buf = torch.zeros(n, device="cuda")
    ref = torch.nn.Parameter(p.clone()); opt = torch.optim.SGD([ref], lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4)
    for t in range(3):
        ref.grad = g.clone()
        opt.step()
        B.sgd_step(p, g, buf, 0.1, 0.9, 1e-4, True, t == 0)
    check(p, ref.detach(), 1e-5, "sgd")


def test_errors_are_loud(B):
    with pytest.raises(_lib.CervixError):
        B.to_nhwc(torch.zeros(1, 3, 4, 4), torch.float32)          # CPU tensor: no fallback
    x = torch.zeros(1, 4, 4, 12, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.CervixError):                           # 12 channels: not 16-byte rows
        B.bn_forward(x, None, torch.ones(12, device="cuda"), torch.zeros(12, device="cuda"),
                     torch.zeros(12, device="cuda"), torch.ones(12, device="cuda"), 0, True, 0.1, 1e-5)
