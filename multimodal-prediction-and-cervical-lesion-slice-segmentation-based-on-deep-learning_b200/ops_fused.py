"""Fused training path of the Xception separable-conv blocks (reference: nets/xception.py:9-31 SeparableConv2d,
:33-73 Block; activate_first=True: relu -> depthwise 3x3 -> bn1 -> pointwise 1x1 -> bn2).

One ``torch.autograd.Function`` per BLOCK drives the whole chain of three separable convolutions by hand, so that
neither BatchNorm output of a separable conv ever touches HBM:

  * bn1 has no ReLU behind it, so its batch-statistics affine is folded into the pointwise weights
    (``pw_fold``); the statistics of the depthwise output come out of the depthwise kernel's epilogue;
  * bn2's statistics come out of the pointwise GEMM's epilogue, and its affine (+ the next layer's ReLU) is applied
    by the NEXT depthwise kernel while it loads its halo tile; only the block output (bn2 + skip [+ ReLU]) is
    materialised;
  * backward: bn1's reduction is read off the weight-gradient GEMM (sum_pix dz*d = sum_o W*G), its apply is the
    data-gradient GEMM's epilogue; both depthwise gradients and bn2's reduction are ONE pass over (dd, x).

Per separable conv this moves 4 tensor passes forward and 11 backward instead of 10 and 19 (csrc/sepconv.cu).
Numerically it is the same computation as the unfused operators in ops.py (statistics in fp64, biased variance for
normalisation, unbiased into the running buffers); tests/test_fused_block_*.py compare the two.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
from torch.autograd import Function

from . import ops
from .backend import ConvGeom, get_backend

ACT_NONE, ACT_RELU = ops.ACT_NONE, ops.ACT_RELU
PARAMS_PER_SEP = 6   # depthwise.weight, bn1.weight, bn1.bias, pointwise.weight, bn2.weight, bn2.bias


class BnBuffers:
    """Running buffers + hyper-parameters of one nn.BatchNorm2d (not differentiable, updated in place)."""
    __slots__ = ("mean", "var", "momentum", "eps")

    def __init__(self, bn: torch.nn.BatchNorm2d):
        track = bn.track_running_stats and bn.running_mean is not None
        self.mean = bn.running_mean if track else None
        self.var = bn.running_var if track else None
        self.momentum = ops._bn_momentum(bn)
        self.eps = float(bn.eps)


class _Sep:
    """What the backward pass needs from one separable conv."""
    __slots__ = ("w9c", "d", "p", "mean1", "invstd1", "scale1", "wpt", "mean2", "invstd2", "scale2", "shift2")


# bn2's statistics come out of the pointwise GEMM's epilogue (column sums of the staged output chunk, the stats-only
# epilogue variant); CERVIX_STATS_EPILOGUE=0 asks for the separate reduction pass instead.  Measured on B200 for the
# 728-channel middle-flow shape: GEMM 49 -> 58 us with the statistics vs 49 + 18 us for GEMM + reduction pass
# (647 vs 644 img/s for the whole step).
_STATS_IN_EPILOGUE = os.environ.get("CERVIX_STATS_EPILOGUE", "1") != "0"
# bn1's backward is applied by the data-gradient GEMM's epilogue (side tile staged by TMA, eight epilogue warps) instead
# of by the depthwise backward kernel (one more tensor read and an in-place pre-pass there); CERVIX_BN1_IN_DGRAD=0
# restores the latter.  Measured on B200 for the 728-channel middle-flow shape: the GEMM grows from 38 to 58 us, the
# depthwise kernel saves ~25 us; 678 vs 670 img/s for the whole step.
_BN1_IN_DGRAD = os.environ.get("CERVIX_BN1_IN_DGRAD", "1") != "0"


def _sep_forward(B, x, in_scale, in_shift, relu_in, dw_w, g1, b1, pw_w, g2, b2, bn1: BnBuffers, bn2: BnBuffers,
                 stats_in_epilogue: bool = _STATS_IN_EPILOGUE) -> _Sep:
    n, h, w, cin = x.shape
    cout = pw_w.shape[0]
    rows = n * h * w
    s = _Sep()
    gd = ConvGeom(n, h, w, cin, cin, 3, 3, 1, 1, 1)
    gp = ConvGeom(n, h, w, cin, cout, 1, 1, 1, 0, 1)
    s.w9c = B.pack_dw_weight(dw_w.detach())
    s.d, st1 = B.dwf_fwd(x, s.w9c, in_scale, in_shift, relu_in, gd, True)
    s.mean1, s.invstd1, s.scale1, shift1 = B.bn_affine(st1, rows, g1.detach(), b1.detach(), bn1.mean, bn1.var,
                                                        bn1.momentum, bn1.eps)
    wp, s.wpt, bias = B.pw_fold(pw_w.detach(), s.scale1, shift1, x.dtype)
    # bn2 removes any per-channel constant, so the folded bn1 shift (bias) is NOT added to p: the stored tensor is
    # p - bias, bn2's statistics are taken on it, and only bn2's running mean needs the constant back.
    if stats_in_epilogue:
        s.p, st2 = B.conv_fwd_ex(s.d, wp, None, gp, None, None, True)
    else:
        s.p = B.conv_fwd(s.d, wp, None, gp, True)
        st2 = B.bn_stats(s.p)
    s.mean2, s.invstd2, s.scale2, s.shift2 = B.bn_affine(st2, rows, g2.detach(), b2.detach(), bn2.mean, bn2.var,
                                                          bn2.momentum, bn2.eps, bias)
    return s


def _sep_backward(B, s: _Sep, dp, x, in_scale, in_shift, relu_in, pw_w, addend, want_sums):
    """dp: gradient w.r.t. the pointwise output p.  Returns (gx, sums, d_dw, d_g1, d_b1, d_pw) where gx is the
    gradient w.r.t. the (virtual) pre-ReLU input in_scale*x+in_shift and sums = (sum gx, sum gx*x)."""
    n, h, w, cin = x.shape
    cout = pw_w.shape[0]
    rows = n * h * w
    gd = ConvGeom(n, h, w, cin, cin, 3, 3, 1, 1, 1)
    gp = ConvGeom(n, h, w, cin, cout, 1, 1, 1, 0, 1)
    G = B.conv_wgrad(s.d, dp, gp, True)
    d_pw, d_g1, d_b1, negk, kmean = B.pw_bwd_coef(G, pw_w.detach(), s.scale1, s.invstd1, s.mean1, rows)
    if _BN1_IN_DGRAD:
        # bn1's backward applied by the data-gradient GEMM's epilogue from the fp32 accumulator:
        #   dd = dp W^T (scaled by bn1) + negk (.) d + kmean ; the depthwise kernel then needs neither d nor a pre-pass
        dd = B.conv_dgrad_ex(dp, s.wpt, gp, kmean, s.d, negk)
        gx, dw9c, sums = B.dwf_bwd(dd, None, None, None, x, s.w9c, in_scale, in_shift, relu_in, addend, gd, want_sums)
    else:
        e = B.conv_dgrad(dp, s.wpt, gp, True)      # = scale1 (.) dz ; bn1's backward is applied by dwf_bwd on load
        gx, dw9c, sums = B.dwf_bwd(e, s.d, negk, kmean, x, s.w9c, in_scale, in_shift, relu_in, addend, gd, want_sums)
    return gx, sums, B.unpack_dw_wgrad(dw9c), d_g1, d_b1, d_pw


class SepChainFn(Function):
    """A chain of K fused separable convs followed by ONE materialisation:
         out = act( bn2_K(p_K) + res ),   res = the chain input itself (identity-skip blocks 4-19), a separate tensor
         (the skip-conv branch of blocks with a 1x1 skip, xception.py:43-47,72) or nothing (entry-flow blocks, whose
         strided third separable conv runs on the operator path)."""

    @staticmethod
    def forward(ctx, inp, residual, res_is_inp: bool, act: int, buffers: Sequence[BnBuffers], chain, *params):
        B = get_backend()
        inp = inp.contiguous()
        ctx.chain = chain          # ops.GradChain shared with the block's 1x1 skip conv (both consume the block input)
        K = len(params) // PARAMS_PER_SEP
        x, sc, sh = inp, None, None
        seps: List[_Sep] = []
        for k in range(K):
            pk = params[k * PARAMS_PER_SEP:(k + 1) * PARAMS_PER_SEP]
            s = _sep_forward(B, x, sc, sh, True, *pk, buffers[2 * k], buffers[2 * k + 1])
            seps.append(s)
            x, sc, sh = s.p, s.scale2, s.shift2
        res = inp if res_is_inp else (None if residual is None else residual.contiguous())
        out = B.affine_act(x, sc, sh, res, act)
        # inputs and the OUTPUT go through save_for_backward (an output kept as a plain attribute would close the cycle
        # out -> grad_fn -> ctx -> out, which only the cyclic GC frees, and would escape autograd's in-place checks);
        # the intermediates of the chain are not graph tensors and simply live on ctx until the graph is freed
        ctx.save_for_backward(inp, out if act != ACT_NONE else None, *params)
        ctx.seps, ctx.act = seps, act
        ctx.has_res, ctx.res_is_inp = res is not None, res_is_inp
        return out

    @staticmethod
    def backward(ctx, dout):
        B = get_backend()
        seps, act = ctx.seps, ctx.act
        inp, out, *params = ctx.saved_tensors
        K = len(seps)
        dout = dout.contiguous()
        rows = inp.numel() // inp.shape[-1]
        grads: List[Optional[torch.Tensor]] = [None] * (K * PARAMS_PER_SEP)
        s = seps[K - 1]
        lbase = (K - 1) * PARAMS_PER_SEP
        sums = B.bn_bwd_sums(dout, out, s.p, act)
        a, b, cc, dgam, dbet = B.bn_bwd_coef(sums, rows, s.mean2, s.invstd2, params[lbase + 4].detach())
        dp, gres = B.bn_bwd_affine(dout, out, s.p, a, b, cc, act, ctx.has_res and act != ACT_NONE)
        if ctx.has_res and gres is None:
            gres = dout                       # no activation behind the add: the skip branch receives dout itself
        grads[lbase + 4], grads[lbase + 5] = dgam, dbet
        dinp = None
        for k in range(K - 1, -1, -1):
            s = seps[k]
            base = k * PARAMS_PER_SEP
            if k > 0:
                prev = seps[k - 1]
                x, sc, sh = prev.p, prev.scale2, prev.shift2
            else:
                x, sc, sh = inp, None, None
            addend = None
            if k == 0:
                # what else flows into the block input is added by the depthwise backward kernel itself: the identity
                # skip's gradient, or the running sum of the other consumers' gradients (the 1x1 skip conv's, if its
                # backward has already run)
                addend = gres if ctx.res_is_inp else (ctx.chain.acc if ctx.chain is not None else None)
            gx, sums, d_dw, d_g1, d_b1, d_pw = _sep_backward(B, s, dp, x, sc, sh, True, params[base + 3], addend, k > 0)
            grads[base + 0], grads[base + 1], grads[base + 2], grads[base + 3] = d_dw, d_g1, d_b1, d_pw
            if k > 0:
                pbase = (k - 1) * PARAMS_PER_SEP
                a, b, cc, dgam, dbet = B.bn_bwd_coef(sums, rows, prev.mean2, prev.invstd2, params[pbase + 4].detach())
                dp, _ = B.bn_bwd_affine(gx, None, prev.p, a, b, cc, ACT_NONE, False)
                grads[pbase + 4], grads[pbase + 5] = dgam, dbet
            else:
                dinp = gx if ctx.chain is None else ctx.chain.take(lambda _sum_already_inside: gx)
        dres = gres if (ctx.has_res and not ctx.res_is_inp) else None
        return (dinp, dres, None, None, None, None, *grads)


def sep_params(sep) -> list:
    return [sep.depthwise.weight, sep.bn1.weight, sep.bn1.bias, sep.pointwise.weight, sep.bn2.weight, sep.bn2.bias]


def chain_fusable(seps, inp: torch.Tensor) -> bool:
    """Training-mode separable convs (activate_first=True) that the fused kernels cover."""
    if not torch.is_grad_enabled():
        return False
    B = get_backend()
    ok = getattr(B, "sepconv_fused_ok", None)
    if ok is None:
        return False
    for sep in seps:
        dw = sep.depthwise
        if not (sep.activate_first and sep.bn1.training and sep.bn2.training and sep.bn1.affine and sep.bn2.affine):
            return False
        if not ok(inp, dw.in_channels, sep.pointwise.out_channels, dw.stride[0], dw.dilation[0], dw.padding[0]):
            return False
    return True


def identity_block_fusable(block, inp: torch.Tensor) -> bool:
    return block.skip is None and chain_fusable((block.sepconv1, block.sepconv2, block.sepconv3), inp)


def sep_chain(seps, inp: torch.Tensor, residual: Optional[torch.Tensor], res_is_inp: bool, act: int, chain=None) -> torch.Tensor:
    buffers, params = [], []
    for sep in seps:
        buffers += [BnBuffers(sep.bn1), BnBuffers(sep.bn2)]
        params += sep_params(sep)
        for bn in (sep.bn1, sep.bn2):
            if bn.track_running_stats and bn.num_batches_tracked is not None and not ops._DEFER_NBT[0]:
                bn.num_batches_tracked.add_(1)
    return SepChainFn.apply(inp, residual, res_is_inp, act, buffers, chain, *params)


def identity_block(block, inp: torch.Tensor, relu_out: bool) -> torch.Tensor:
    return sep_chain((block.sepconv1, block.sepconv2, block.sepconv3), inp, None, True,
                     ACT_RELU if relu_out else ACT_NONE)
