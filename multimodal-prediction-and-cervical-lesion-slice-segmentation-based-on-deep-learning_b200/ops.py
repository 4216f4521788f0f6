"""Autograd operators of the B200 DeepLabv3+ path.

Each ``torch.autograd.Function`` below is a thin host-side shell: forward and backward call
the CUDA kernels through the C ABI (backend.py) and nothing else.  Activations flowing between
operators are NHWC tensors in the engine's compute dtype (bf16 for training, fp32 for the
exact-parity mode); parameters stay in the reference's fp32 OIHW/`nn.BatchNorm2d` layout so the
reference's ``state_dict``/optimizers/DDP see ordinary ``nn.Parameter``s.
"""
from __future__ import annotations

import os

import torch
from torch.autograd import Function

from . import _lib
from .backend import ConvGeom, get_backend

ACT_NONE, ACT_RELU, ACT_RELU6 = _lib.ACT_NONE, _lib.ACT_RELU, _lib.ACT_RELU6


_TC_WARNED = [False]


def _tc_ok(x: torch.Tensor, cin: int, cout: int) -> bool:
    """Tensor-core (tcgen05) eligibility of a dense conv: bf16 storage, 16-byte rows.  There is no quiet slow path:
    bf16 convolutions on a device that is not sm_100 raise, and the debugging switch ``CERVIX_DISABLE_TC=1`` (the SIMT
    kernels, ~10x slower) announces itself on stderr."""
    B = get_backend()
    if x.dtype != torch.bfloat16 or cin % 8 != 0 or cout % 8 != 0:
        return False
    if os.environ.get("CERVIX_DISABLE_TC") == "1":
        if not _TC_WARNED[0]:
            _TC_WARNED[0] = True
            import sys
            print("cervix_b200: CERVIX_DISABLE_TC=1 - bf16 convolutions run on the SIMT kernels (debugging mode, ~10x slower)",
                  file=sys.stderr, flush=True)
        return False
    probe = getattr(B, "is_sm100", None)
    if probe is None:               # the CPU emulation backend of the host-logic tests
        return False
    if not probe():
        raise RuntimeError("cervix_b200: the bf16 convolution path needs an sm_100 device (tcgen05); "
                           "there is no fallback - use set_compute_dtype(torch.float32) for the exact SIMT path")
    return True


class ToNHWC(Function):
    """NCHW fp32 image batch -> NHWC compute dtype (entry of DeepLab.forward)."""

    @staticmethod
    def forward(ctx, x, dtype):
        return get_backend().to_nhwc(x.contiguous().float(), dtype)

    @staticmethod
    def backward(ctx, dy):
        return get_backend().to_nchw(dy.contiguous()), None


class ToNCHW(Function):
    @staticmethod
    def forward(ctx, x):
        ctx.dtype = x.dtype
        return get_backend().to_nchw(x.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return get_backend().to_nhwc(dy.contiguous(), ctx.dtype)


class GradChain:
    """Gradient accumulation of a tensor with several consumers WITHOUT separate add passes (ASPP: five consumers of the
    [B,32,32,2048] backbone output, deeplabv3_plus.py:89-114 - autograd would sum their gradients with four ATen adds of
    134 MB tensors each).  The consumers share one chain: each backward adds its gradient to the running sum - the
    tensor-core data gradient takes the sum as the epilogue's side input (cvx_conv_dgrad_tc_ex), so the add rides a pass
    that exists anyway - and returns None to autograd until the last consumer hands over the total.  Used with ``fanout``,
    which gives every consumer its own alias of the tensor."""

    def __init__(self, consumers: int):
        self.n = self.left = consumers
        self.acc = None

    def take(self, make):
        """``make(side)`` -> this consumer's gradient with ``side`` (the running sum, or None) already added."""
        self.acc = make(self.acc)
        return self._emit()

    def add(self, dx):
        """A consumer that cannot add inside its own kernel."""
        self.acc = dx if self.acc is None else self.acc + dx
        return self._emit()

    def _emit(self):
        self.left -= 1
        if self.left:
            return None
        out, self.acc, self.left = self.acc, None, self.n
        return out


class Fanout(Function):
    """k aliases of one tensor; the backward expects at most one real gradient per chain (see GradChain) and falls back to
    summing whatever arrives."""

    @staticmethod
    def forward(ctx, x, k):
        ctx.set_materialize_grads(False)      # consumers that handed their gradient to the chain return None: keep it None
        return tuple(x.view_as(x) for _ in range(k))

    @staticmethod
    def backward(ctx, *grads):
        live = [g for g in grads if g is not None]
        if not live:
            return None, None
        total = live[0]
        for g in live[1:]:
            total = total + g
        return total, None


_ONES = {}


def _ones_f32(n, device):
    key = (n, str(device))
    if key not in _ONES:
        _ONES[key] = torch.ones(n, dtype=torch.float32, device=device)
    return _ONES[key]


def _chained_dgrad(chain, B, dz, wpt, g, tc):
    """Data gradient of a dense conv, added to the chain's running sum inside the GEMM epilogue when the layer runs on
    the tensor-core path in bf16 (the side input is bf16); plain add otherwise."""
    if chain is None:
        return B.conv_dgrad(dz, wpt, g, tc)
    if tc and dz.dtype == torch.bfloat16 and hasattr(B, "conv_dgrad_ex"):
        return chain.take(lambda side: B.conv_dgrad(dz, wpt, g, tc) if side is None
                          else B.conv_dgrad_ex(dz, wpt, g, None, side.contiguous(), _ones_f32(g.cin, dz.device)))
    return chain.add(B.conv_dgrad(dz, wpt, g, tc))


class Conv2d(Function):
    """Dense nn.Conv2d (groups=1) on NHWC.  weight: fp32 OIHW parameter; bias: fp32 or None."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, dil, chain=None):
        B = get_backend()
        x = x.contiguous()
        n, h, w, cin = x.shape
        cout, cin_w, kh, kw = weight.shape
        assert cin_w == cin, "conv: channel mismatch %d vs %d" % (cin_w, cin)
        ctx.chain = chain
        tc = _tc_ok(x, cin, cout)
        sub = 1
        if tc and stride != 1 and kh == 1 and kw == 1 and pad == 0:
            # 1x1 stride-s conv == 1x1 stride-1 conv of the subsampled input
            x = B.subsample(x, stride)
            sub, stride = stride, 1
        g = ConvGeom(x.shape[0], x.shape[1], x.shape[2], cin, cout, kh, kw, stride, pad, dil)
        # strided kxk convs: the tensor-core forward takes stride 2 (TMA element strides); their
        # gradients use the generic kernels
        tc_fwd = tc and stride in (1, 2)
        tc_bwd = tc and stride == 1
        wp = B.pack_weight(weight.detach(), x.dtype, False)
        y = B.conv_fwd(x, wp, None if bias is None else bias.detach(), g, tc_fwd)
        tc = tc_bwd
        ctx.save_for_backward(x, weight)
        ctx.g, ctx.tc, ctx.sub, ctx.in_hw, ctx.has_bias = g, tc, sub, (h, w), bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        B = get_backend()
        x, weight = ctx.saved_tensors
        g, tc = ctx.g, ctx.tc
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wpt = B.pack_weight(weight.detach(), dy.dtype, True)
            if ctx.sub != 1:
                dx = B.subsample_bwd(B.conv_dgrad(dy, wpt, g, tc), ctx.in_hw[0], ctx.in_hw[1], ctx.sub)
                if ctx.chain is not None:
                    dx = ctx.chain.add(dx)
            else:
                dx = _chained_dgrad(ctx.chain, B, dy, wpt, g, tc)
        if ctx.needs_input_grad[1]:
            dwp = B.conv_wgrad(x, dy, g, tc)
            dw = B.unpack_wgrad(dwp, g.cout, g.cin, g.kh, g.kw)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = B.bias_grad(dy)
        return dx, dw, db, None, None, None, None


class ConvBnAct(Function):
    """Training-mode ``act(BatchNorm2d(conv(x)) + residual)`` for a bias-free dense conv on the tensor-core path, with
    the batch statistics taken in the GEMM epilogue (cvx_conv_fwd_tc_ex) instead of by a separate pass over the conv
    output: conv -> (sum, sum^2 per channel) -> scale/shift (cvx_bn_affine, also updates the running buffers) ->
    one apply pass.  Same math as ``Conv2d`` followed by ``BatchNormAct``."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, residual, stride, pad, dil, momentum, eps,
                act, chain=None):
        B = get_backend()
        ctx.chain = chain
        x = x.contiguous()
        n, h, w, cin = x.shape
        cout, _, kh, kw = weight.shape
        sub = 1
        if stride != 1 and kh == 1 and kw == 1 and pad == 0:
            x = B.subsample(x, stride)
            sub, stride = stride, 1
        g = ConvGeom(x.shape[0], x.shape[1], x.shape[2], cin, cout, kh, kw, stride, pad, dil)
        wp = B.pack_weight(weight.detach(), x.dtype, False)
        p, stats = B.conv_fwd_ex(x, wp, None if bias is None else bias.detach(), g, None, None, True)
        rows = p.numel() // cout
        mean, invstd, scale, shift = B.bn_affine(stats, rows, gamma.detach(), beta.detach(), running_mean, running_var,
                                                 momentum, eps)
        if residual is not None:
            residual = residual.contiguous()
        y = B.affine_act(p, scale, shift, residual, act)
        mask_from_x = act != ACT_NONE and residual is None
        ctx.save_for_backward(x, weight, p, y if (act != ACT_NONE and not mask_from_x) else None, gamma, mean, invstd,
                              beta if mask_from_x else None)
        ctx.g, ctx.sub, ctx.in_hw, ctx.act, ctx.has_res = g, sub, (h, w), act, residual is not None
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        B = get_backend()
        x, weight, p, y, gamma, mean, invstd, beta = ctx.saved_tensors
        g = ctx.g
        dz, dres, dgamma, dbeta = B.bn_backward(dy.contiguous(), p, y, gamma.detach(), mean, invstd, ctx.act, True,
                                                ctx.has_res and ctx.needs_input_grad[7],
                                                None if beta is None else beta.detach())
        tc = g.stride == 1
        dx = dw = None
        if ctx.needs_input_grad[0]:
            wpt = B.pack_weight(weight.detach(), dz.dtype, True)
            if ctx.sub != 1:
                dx = B.subsample_bwd(B.conv_dgrad(dz, wpt, g, tc), ctx.in_hw[0], ctx.in_hw[1], ctx.sub)
                if ctx.chain is not None:
                    dx = ctx.chain.add(dx)
            else:
                dx = _chained_dgrad(ctx.chain, B, dz, wpt, g, tc)
        if ctx.needs_input_grad[1]:
            dw = B.unpack_wgrad(B.conv_wgrad(x, dz, g, tc), g.cout, g.cin, g.kh, g.kw)
        db = B.bias_grad(dz) if (ctx.has_bias and ctx.needs_input_grad[2]) else None   # (sums to ~0 behind a BatchNorm)
        return (dx, dw, db, dgamma if ctx.needs_input_grad[3] else None, dbeta if ctx.needs_input_grad[4] else None,
                None, None, dres, None, None, None, None, None, None, None)


class MaxPool3x3S2(Function):
    """nn.MaxPool2d(kernel_size=3, stride=2, padding=1) on NHWC."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        y = get_backend().maxpool_fwd(x)
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        return get_backend().maxpool_bwd(x, y, dy.contiguous())


def conv2d_narrow_in(x, weight, stride, pad):
    """Inference-only dense conv for C_in <= 4 and a large filter (the 7x7/2 ResNet stem): im2col into a
    K-padded patch tensor, then a 1x1 conv on the tensor cores.  Falls back to ``conv2d`` when a gradient
    is needed or the engine runs in fp32."""
    cout, cin, kh, kw = weight.shape
    if (torch.is_grad_enabled() and (weight.requires_grad or x.requires_grad)) or not _tc_ok(x, 8, cout):
        return conv2d(x, weight, None, stride, pad, 1)
    B = get_backend()
    x = x.contiguous()
    n, h, w, _ = x.shape
    g = ConvGeom(n, h, w, cin, cout, kh, kw, stride, pad, 1)
    k = kh * kw * cin
    kpad = (k + 7) // 8 * 8
    patches = B.im2col_narrow(x, g, kpad)
    # [cout, cin, kh, kw] -> [cout, (tap, ci)] zero-padded to kpad, as a 1x1 filter
    w2 = weight.detach().permute(0, 2, 3, 1).reshape(cout, k)
    w2 = torch.nn.functional.pad(w2, (0, kpad - k)).reshape(cout, kpad, 1, 1).contiguous()
    return conv2d(patches, w2, None, 1, 0, 1)


class DepthwiseConv3x3(Function):
    """nn.Conv2d(C, C, 3, groups=C, bias=False) on NHWC with an optional fused input ReLU."""

    @staticmethod
    def forward(ctx, x, weight, stride, pad, dil, relu_in):
        B = get_backend()
        x = x.contiguous()
        n, h, w, c = x.shape
        assert tuple(weight.shape) == (c, 1, 3, 3)
        g = ConvGeom(n, h, w, c, c, 3, 3, stride, pad, dil)
        w9c = B.pack_dw_weight(weight.detach())
        y = B.dw_fwd(x, w9c, g, relu_in)
        ctx.save_for_backward(x, w9c)
        ctx.g, ctx.relu_in = g, relu_in
        return y

    @staticmethod
    def backward(ctx, dy):
        B = get_backend()
        x, w9c = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = B.dw_bwd_data(dy, w9c, x, ctx.g, ctx.relu_in)
        if ctx.needs_input_grad[1]:
            dw = B.unpack_dw_wgrad(B.dw_bwd_weight(x, dy, ctx.g, ctx.relu_in))
        return dx, dw, None, None, None, None


class BatchNormAct(Function):
    """y = act(BatchNorm2d(x) + residual) with nn.BatchNorm2d training/eval semantics."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, residual, training, momentum, eps, act):
        B = get_backend()
        x = x.contiguous()
        if residual is not None:
            residual = residual.contiguous()
        y, mean, invstd = B.bn_forward(x, residual, gamma.detach(), beta.detach(), running_mean, running_var, act,
                                       training, momentum, eps)
        # without a residual the backward recomputes the activation mask from x (saves two passes over y)
        ctx.mask_from_x = act != ACT_NONE and residual is None
        ctx.save_for_backward(x, y if (act != ACT_NONE and not ctx.mask_from_x) else None, gamma, mean, invstd,
                              beta if ctx.mask_from_x else None)
        ctx.act, ctx.training, ctx.has_res = act, training, residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        B = get_backend()
        x, y, gamma, mean, invstd, beta = ctx.saved_tensors
        dx, dres, dgamma, dbeta = B.bn_backward(dy.contiguous(), x, y, gamma.detach(), mean, invstd, ctx.act,
                                                ctx.training, ctx.has_res and ctx.needs_input_grad[5],
                                                None if beta is None else beta.detach())
        if not ctx.needs_input_grad[1]:
            dgamma = None
        if not ctx.needs_input_grad[2]:
            dbeta = None
        return dx, dgamma, dbeta, None, None, dres, None, None, None, None


class ReLU(Function):
    @staticmethod
    def forward(ctx, x):
        y = get_backend().relu_fwd(x.contiguous())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return get_backend().relu_bwd(dy.contiguous(), y)


class Add(Function):
    @staticmethod
    def forward(ctx, a, b):
        return get_backend().add(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


class CatChannels(Function):
    """torch.cat(dim=1) of the reference == channel concat on NHWC."""

    @staticmethod
    def forward(ctx, *xs):
        ctx.widths = [int(x.shape[3]) for x in xs]
        return get_backend().cat_channels([x.contiguous() for x in xs])

    @staticmethod
    def backward(ctx, dy):
        B = get_backend()
        dy = dy.contiguous()
        outs, off = [], 0
        for i, c in enumerate(ctx.widths):
            outs.append(B.slice_channels(dy, off, c) if ctx.needs_input_grad[i] else None)
            off += c
        return tuple(outs)


class GlobalAvgPool(Function):
    """torch.mean over H then W with keepdim (ASPP branch 5)."""

    @staticmethod
    def forward(ctx, x, chain=None):
        n, h, w, c = x.shape
        ctx.hw = (h, w)
        ctx.chain = chain
        return get_backend().spatial_reduce(x.contiguous(), 1.0 / (h * w))

    @staticmethod
    def backward(ctx, dy):
        h, w = ctx.hw
        dx = get_backend().spatial_broadcast(dy.contiguous(), h, w, 1.0 / (h * w))
        return (dx if ctx.chain is None else ctx.chain.add(dx)), None


class BroadcastHW(Function):
    """F.interpolate(1x1 -> HxW, bilinear, align_corners=True) == broadcast."""

    @staticmethod
    def forward(ctx, x, h, w):
        return get_backend().spatial_broadcast(x.contiguous(), h, w, 1.0)

    @staticmethod
    def backward(ctx, dy):
        return get_backend().spatial_reduce(dy.contiguous(), 1.0), None, None


class UpsampleBilinear(Function):
    """F.interpolate(mode='bilinear', align_corners=True) on NHWC."""

    @staticmethod
    def forward(ctx, x, ho, wo):
        ctx.in_hw = (int(x.shape[1]), int(x.shape[2]))
        return get_backend().upsample_fwd(x.contiguous(), ho, wo)

    @staticmethod
    def backward(ctx, dy):
        return get_backend().upsample_bwd(dy.contiguous(), *ctx.in_hw), None, None


class UpsampleConcat(Function):
    """cat([F.interpolate(x, size of tail, bilinear, align_corners=True), tail], channels) - the decoder's junction of the
    ASPP output with the low-level feature (deeplabv3_plus.py:184-185) - with the upsample writing / its gradient reading
    the concat buffer directly (no upsampled intermediate, no slice copy of the big half in backward)."""

    @staticmethod
    def forward(ctx, x, tail):
        ctx.in_hw, ctx.c = (x.shape[1], x.shape[2]), x.shape[3]
        return get_backend().upsample_concat(x.contiguous(), tail.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return get_backend().upsample_concat_bwd(dy.contiguous(), ctx.c, ctx.in_hw[0], ctx.in_hw[1])


class UpsampleToNCHW(Function):
    """Final F.interpolate of the logits fused with the NHWC -> NCHW fp32 conversion."""

    @staticmethod
    def forward(ctx, x, ho, wo):
        ctx.in_hw, ctx.dtype = (int(x.shape[1]), int(x.shape[2])), x.dtype
        return get_backend().upsample_to_nchw_fwd(x.contiguous(), ho, wo)

    @staticmethod
    def backward(ctx, dy):
        return get_backend().upsample_to_nchw_bwd(dy.contiguous().float(), ctx.in_hw[0], ctx.in_hw[1], ctx.dtype), None, None


_MASKED_DROPOUT = os.environ.get("CERVIX_DROPOUT_MASK", "0") == "1"     # A/B switch: keep the stored byte mask


class Dropout(Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        B = get_backend()
        ctx.p, ctx.seed, ctx.step_dev = p, seed, _STEP_DEV[0]
        ctx.seeded = hasattr(B, "dropout_bwd_seeded") and not _MASKED_DROPOUT   # decisions = f(seed, step, index)
        y, mask = B.dropout_fwd(x.contiguous(), p, seed, _STEP_DEV[0], want_mask=not ctx.seeded) if ctx.seeded \
            else B.dropout_fwd(x.contiguous(), p, seed, _STEP_DEV[0])
        ctx.save_for_backward(mask)
        return y

    @staticmethod
    def backward(ctx, dy):
        B = get_backend()
        if ctx.seeded:       # the step counter is read on the device: same value as in this step's forward
            return B.dropout_bwd_seeded(dy.contiguous(), ctx.p, ctx.seed, ctx.step_dev), None, None
        (mask,) = ctx.saved_tensors
        return B.dropout_bwd(dy.contiguous(), mask, ctx.p), None, None


class SegLoss(Function):
    """Weighted sum of the reference's CE / focal / dice losses plus the f_score metric, from
    one statistics pass over fp32 NCHW logits.  Returns a 4-vector (ce, focal, dice, f_score);
    gradients flow from the first three."""

    @staticmethod
    def forward(ctx, logits, target, onehot, cls_w, alpha, gamma, beta, smooth, thr):
        B = get_backend()
        logits = logits.contiguous()
        target = target.contiguous()
        if onehot is not None:
            onehot = onehot.contiguous().float()
        if cls_w is not None:
            cls_w = cls_w.contiguous().float()
        c = int(logits.shape[1])
        stats = B.seg_loss_stats(logits, target, onehot, cls_w, alpha, gamma, thr)
        res = B.seg_loss_finalize(stats, c, beta, smooth)
        ctx.save_for_backward(logits, target, onehot, cls_w, stats)
        ctx.cfg = (alpha, gamma, beta, smooth)
        return res

    @staticmethod
    def backward(ctx, dres):
        logits, target, onehot, cls_w, stats = ctx.saved_tensors
        alpha, gamma, beta, smooth = ctx.cfg
        g = dres.contiguous().float()[:3].contiguous()
        d = get_backend().seg_loss_grad(logits, target, onehot, cls_w, stats, g, alpha, gamma, beta, smooth)
        return d, None, None, None, None, None, None, None, None


# ---------------------------------------------------------------------------- functional API
def to_nhwc(x, dtype):
    return ToNHWC.apply(x, dtype)


# ---- cut points: tensors at which a trainer may split the backward pass into two phases (engine.SegTrainer, data
# parallel: the all-reduce of the gradients produced by the first phase runs under the second phase).  When a sink is set,
# the consumer side continues from a detached leaf copy (a view, no kernel), so the autograd graph falls apart into the
# part after the cut (loss -> proxies, late parameters) and the part before it (original tensors -> early parameters);
# the trainer feeds the proxies' gradients into the originals.  A no-op otherwise.
_CUT_SINK = [None]


def cut_point(x, tag: str):
    sink = _CUT_SINK[0]
    if sink is None or not torch.is_grad_enabled() or not x.requires_grad:
        return x
    proxy = x.detach().requires_grad_(True)
    sink[tag] = (x, proxy)
    return proxy


def conv2d(x, weight, bias=None, stride=1, pad=0, dil=1, chain=None):
    return Conv2d.apply(x, weight, bias, stride, pad, dil, chain)


def fanout(x, k: int):
    """k aliases of ``x`` for k consumers that share a ``GradChain``."""
    return Fanout.apply(x, k)


def dwconv3x3(x, weight, stride=1, pad=1, dil=1, relu_in=False):
    return DepthwiseConv3x3.apply(x, weight, stride, pad, dil, relu_in)


# When a trainer owns the step it bumps every ``num_batches_tracked`` with ONE foreach launch
# (engine.SegTrainer) instead of 141 one-element kernels; see defer_batch_counters().
_DEFER_NBT = [False]


class defer_batch_counters:
    def __enter__(self):
        self.prev = _DEFER_NBT[0]
        _DEFER_NBT[0] = True

    def __exit__(self, *exc):
        _DEFER_NBT[0] = self.prev


def _bn_momentum(bn) -> float:
    """The reference only builds BatchNorm2d(momentum=0.0003 | 0.1) (xception.py:7, deeplabv3_plus.py:64).  ``momentum=None``
    means a cumulative moving average in torch, which the kernels do not implement: refuse instead of silently freezing
    the running statistics."""
    if bn.momentum is None:
        raise ValueError("cervix_b200: BatchNorm2d(momentum=None) (cumulative average) is not supported")
    return float(bn.momentum)


def batchnorm_act(x, bn: torch.nn.BatchNorm2d, act=ACT_NONE, residual=None):
    """Apply an ``nn.BatchNorm2d`` parameter holder to an NHWC tensor (never calls bn.forward)."""
    training = bn.training or (bn.running_mean is None)
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None and not _DEFER_NBT[0]:
        bn.num_batches_tracked.add_(1)
    momentum = _bn_momentum(bn)
    return BatchNormAct.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual, training, momentum,
                              bn.eps, act)


# Measured (batch 32, A/B over the whole step): the composite is a wash - 45.8 ms with it off, 45.7 .. 46.2 ms with it on
# for the long-K layers only (K >= 1024: decoder 3x3, ASPP), 46.1 ms for all layers - because the epilogue's column sums
# cost about what the separate statistics pass does.  It therefore stays OFF by default; CERVIX_CONV_BN_MINK=<K> turns
# it on for convolutions whose reduction length cin*kh*kw is at least K.
_CONV_BN_MIN_K = int(os.environ.get("CERVIX_CONV_BN_MINK", str(1 << 30)))


def conv_bn_act(x, conv_weight, stride, pad, dil, bn: torch.nn.BatchNorm2d, act=ACT_NONE, residual=None, bias=None,
                chain=None):
    """``act(bn(conv(x)) + residual)``: one composite with the BatchNorm statistics from the GEMM
    epilogue when the layer trains on the tensor-core path, else ``conv2d`` followed by ``batchnorm_act``."""
    cout, cin, kh, kw = conv_weight.shape
    B = get_backend()
    fused = (bn.training and bn.running_mean is not None and _tc_ok(x, cin, cout) and hasattr(B, "conv_fwd_ex")
             and cin * kh * kw >= _CONV_BN_MIN_K
             and (stride == 1 or (kh == 1 and kw == 1 and pad == 0) or stride == 2)
             and x.shape[0] * x.shape[1] * x.shape[2] >= 1024)
    if not fused:
        return batchnorm_act(conv2d(x, conv_weight, bias, stride, pad, dil, chain), bn, act, residual)
    if bn.track_running_stats and bn.num_batches_tracked is not None and not _DEFER_NBT[0]:
        bn.num_batches_tracked.add_(1)
    momentum = _bn_momentum(bn)
    return ConvBnAct.apply(x, conv_weight, bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual, stride,
                           pad, dil, momentum, bn.eps, act, chain)


def relu(x):
    return ReLU.apply(x)


def maxpool3x3s2(x):
    return MaxPool3x3S2.apply(x)


def cat_channels(xs):
    return CatChannels.apply(*xs)


def global_avg_pool(x, chain=None):
    return GlobalAvgPool.apply(x, chain)


def broadcast_hw(x, h, w):
    return BroadcastHW.apply(x, h, w)


def upsample_bilinear(x, ho, wo):
    if x.shape[1] == ho and x.shape[2] == wo:
        return x
    return UpsampleBilinear.apply(x, ho, wo)


_NO_UPCAT = os.environ.get("CERVIX_UPSAMPLE_CONCAT", "1") == "0"      # A/B switch


def upsample_concat(x, tail):
    B = get_backend()
    vec = 4 if x.dtype == torch.float32 else 8
    if not hasattr(B, "upsample_concat") or x.shape[3] % vec or tail.shape[3] % vec or _NO_UPCAT:
        return cat_channels([upsample_bilinear(x, tail.shape[1], tail.shape[2]), tail])
    return UpsampleConcat.apply(x, tail)


def upsample_to_nchw(x, ho, wo):
    return UpsampleToNCHW.apply(x, ho, wo)


_dropout_counter = [0]
# device int32 step counter installed by a trainer that replays a captured CUDA graph: mixed into every
# dropout seed on the device so each replay draws fresh masks (host-side seeds are frozen in the graph)
_STEP_DEV = [None]


def set_step_counter(t):
    _STEP_DEV[0] = t



def dropout(x, p, training):
    if not training or p <= 0.0:
        return x
    # seed derived from torch's generator so torch.manual_seed governs it (host-side only)
    _dropout_counter[0] += 1
    seed = (torch.initial_seed() * 1000003 + _dropout_counter[0]) & (2 ** 63 - 1)
    return Dropout.apply(x, float(p), seed)


def seg_losses(logits, target, onehot=None, cls_weights=None, alpha=0.5, gamma=2.0, beta=1.0, smooth=1e-5,
               threshold=0.5):
    """(ce, focal, dice, f_score) as a 4-element fp32 tensor."""
    a = 1.0 if alpha is None else float(alpha)
    return SegLoss.apply(logits, target, onehot, cls_weights, a, float(gamma), float(beta), float(smooth),
                         float(threshold))
