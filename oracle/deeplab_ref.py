"""CPU oracle for the DeepLabv3+ hot path (TEST INFRASTRUCTURE ONLY).

This file is a from-scratch *functional* restatement, in plain torch fp32 ops, of the
reference's segmentation model.  It is the checker for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may
import it.  The product package never imports anything under ``oracle/``.

Parity pin: the restatement is validated against the *unmodified* reference modules
(imported from /root/reference in the build container) by ``oracle/make_golden.py``;
the resulting vectors are committed under ``tests/golden/`` and re-checked by
``tests/test_oracle_golden.py`` on every run.  The reference itself ships no tests or
golden vectors, so that live comparison is the pin.

Reference symbols restated here (paths relative to
/root/reference/Segmentation/deeplabv3+/):
  * nets/xception.py:9-31    SeparableConv2d      -> _sepconv
  * nets/xception.py:33-73   Block                -> _xception_block
  * nets/xception.py:76-182  Xception             -> xception_forward
  * nets/mobilenetv2.py:24-72 InvertedResidual    -> _inverted_residual
  * nets/deeplabv3_plus.py:7-49 MobileNetV2 wrapper (stride->dilation rewrite)
                                                  -> mobilenet_plan / mobilenet_forward
  * nets/deeplabv3_plus.py:56-114 ASPP            -> aspp_forward
  * nets/deeplabv3_plus.py:169-188 DeepLab.forward -> deeplab_forward

All functions take a flat ``state`` dict keyed exactly like the reference's
``state_dict()`` (857 entries for Xception, 371 for MobileNetV2).  Tensors in the dict
may require grad, in which case torch autograd supplies the oracle gradients.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]

XCEPTION_BN_MOMENTUM = 0.0003  # nets/xception.py:7
DEFAULT_BN_MOMENTUM = 0.1      # nn.BatchNorm2d default (ASPP, decoder, MobileNetV2)
BN_EPS = 1e-5


# ----------------------------------------------------------------------------- primitives
def _bn(x, state: State, prefix: str, training: bool, momentum: float):
    """nn.BatchNorm2d: batch statistics (biased var for normalisation, unbiased into the
    running buffer) when training, running statistics otherwise."""
    rm = state[prefix + ".running_mean"]
    rv = state[prefix + ".running_var"]
    if training:
        nbt = state.get(prefix + ".num_batches_tracked")
        if nbt is not None:
            nbt += 1
    return F.batch_norm(x, rm, rv, state[prefix + ".weight"], state[prefix + ".bias"],
                        training, momentum, BN_EPS)


def _conv(x, state: State, prefix: str, stride=1, padding=0, dilation=1, groups=1):
    return F.conv2d(x, state[prefix + ".weight"], state.get(prefix + ".bias"),
                    stride, padding, dilation, groups)


# ----------------------------------------------------------------------------- Xception
def _sepconv(x, state, p, stride, dilation, activate_first, training):
    # nets/xception.py:21-31
    c = x.shape[1]
    if activate_first:
        x = F.relu(x)
    x = _conv(x, state, p + ".depthwise", stride, dilation, dilation, groups=c)
    x = _bn(x, state, p + ".bn1", training, XCEPTION_BN_MOMENTUM)
    if not activate_first:
        x = F.relu(x)
    x = _conv(x, state, p + ".pointwise")
    x = _bn(x, state, p + ".bn2", training, XCEPTION_BN_MOMENTUM)
    if not activate_first:
        x = F.relu(x)
    return x


def _xception_block(inp, state, p, strides, atrous, training):
    """nets/xception.py:33-73.  Note the aliasing quirk: when the block has no skip conv
    the first SeparableConv2d's relu0 is in-place (``inplace=self.head_relu``), so the
    identity branch adds relu(inp), not inp.  Returns (out, hook) where hook is the
    *pre-ReLU* output of sepconv2 (only block2 keeps it un-aliased: inplace=False at
    xception.py:104)."""
    has_skip = (p + ".skip.weight") in state
    if has_skip:
        skip = _conv(inp, state, p + ".skip", stride=strides)
        skip = _bn(skip, state, p + ".skipbn", training, XCEPTION_BN_MOMENTUM)
        main_in = inp
    else:
        main_in = F.relu(inp)
        skip = main_in
    x = _sepconv(main_in, state, p + ".sepconv1", 1, atrous[0], True, training)
    x = _sepconv(x, state, p + ".sepconv2", 1, atrous[1], True, training)
    hook = x
    x = _sepconv(x, state, p + ".sepconv3", strides, atrous[2], True, training)
    return x + skip, hook


def xception_forward(x, state: State, downsample_factor: int = 16, training: bool = False,
                     prefix: str = "backbone"):
    if downsample_factor == 8:
        stride_list = [2, 1, 1]
    elif downsample_factor == 16:
        stride_list = [2, 2, 1]
    else:
        # the reference intends ValueError but formats the `os` module -> TypeError
        # (nets/xception.py:94); the product mirrors that, the oracle just refuses.
        raise TypeError("xception: output stride %r is not supported" % (downsample_factor,))
    rate = 16 // downsample_factor
    b = prefix
    x = _conv(x, state, b + ".conv1", 2, 1)
    x = F.relu(_bn(x, state, b + ".bn1", training, XCEPTION_BN_MOMENTUM))
    x = _conv(x, state, b + ".conv2", 1, 1)
    x = F.relu(_bn(x, state, b + ".bn2", training, XCEPTION_BN_MOMENTUM))
    x, _ = _xception_block(x, state, b + ".block1", 2, [1, 1, 1], training)
    x, low = _xception_block(x, state, b + ".block2", stride_list[0], [1, 1, 1], training)
    x, _ = _xception_block(x, state, b + ".block3", stride_list[1], [1, 1, 1], training)
    for i in range(4, 20):
        x, _ = _xception_block(x, state, b + ".block%d" % i, 1, [rate] * 3, training)
    x, _ = _xception_block(x, state, b + ".block20", stride_list[2], [rate] * 3, training)
    x = _sepconv(x, state, b + ".conv3", 1, rate, False, training)
    x = _sepconv(x, state, b + ".conv4", 1, rate, False, training)
    x = _sepconv(x, state, b + ".conv5", 1, rate, False, training)
    return low, x


# ----------------------------------------------------------------------------- MobileNetV2
_MBV2_SETTING = [  # t, c, n, s   (nets/mobilenetv2.py:80-89)
    (1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2),
    (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1),
]


def mobilenet_plan(downsample_factor: int) -> List[dict]:
    """Per-feature-index description after the stride->dilation rewrite of
    nets/deeplabv3_plus.py:18-43.  Index 0 is the stem conv_bn; 1..17 the inverted
    residual blocks (features[:-1] drops the final 1x1)."""
    plan = [dict(kind="stem")]
    inp = 32
    for t, c, n, s in _MBV2_SETTING:
        for i in range(n):
            plan.append(dict(kind="ir", inp=inp, oup=c, stride=s if i == 0 else 1, expand=t,
                             dilation=1, res=(s if i == 0 else 1) == 1 and inp == c))
            inp = c
    down_idx = [2, 4, 7, 14]
    total = len(plan)

    def nostride_dilate(idx, dilate):
        blk = plan[idx]
        if blk["stride"] == 2:
            blk["stride"] = 1
            blk["dilation"] = dilate // 2
        else:
            blk["dilation"] = dilate

    if downsample_factor == 8:
        for i in range(down_idx[-2], down_idx[-1]):
            nostride_dilate(i, 2)
        for i in range(down_idx[-1], total):
            nostride_dilate(i, 4)
    elif downsample_factor == 16:
        for i in range(down_idx[-1], total):
            nostride_dilate(i, 2)
    return plan


def _inverted_residual(x, state, p, blk, training):
    hidden = round(blk["inp"] * blk["expand"])
    d = blk["dilation"]
    y = x
    if blk["expand"] == 1:
        y = _conv(y, state, p + ".conv.0", blk["stride"], d, d, groups=hidden)
        y = F.relu6(_bn(y, state, p + ".conv.1", training, DEFAULT_BN_MOMENTUM))
        y = _conv(y, state, p + ".conv.3")
        y = _bn(y, state, p + ".conv.4", training, DEFAULT_BN_MOMENTUM)
    else:
        y = _conv(y, state, p + ".conv.0")
        y = F.relu6(_bn(y, state, p + ".conv.1", training, DEFAULT_BN_MOMENTUM))
        y = _conv(y, state, p + ".conv.3", blk["stride"], d, d, groups=hidden)
        y = F.relu6(_bn(y, state, p + ".conv.4", training, DEFAULT_BN_MOMENTUM))
        y = _conv(y, state, p + ".conv.6")
        y = _bn(y, state, p + ".conv.7", training, DEFAULT_BN_MOMENTUM)
    return x + y if blk["res"] else y


def mobilenet_forward(x, state: State, downsample_factor: int = 16, training: bool = False,
                      prefix: str = "backbone"):
    plan = mobilenet_plan(downsample_factor)
    f = prefix + ".features"
    x = _conv(x, state, f + ".0.0", 2, 1)
    x = F.relu6(_bn(x, state, f + ".0.1", training, DEFAULT_BN_MOMENTUM))
    low = None
    for i in range(1, len(plan)):
        x = _inverted_residual(x, state, f + ".%d" % i, plan[i], training)
        if i == 3:
            low = x
    return low, x


# ----------------------------------------------------------------------------- ASPP + decoder
def aspp_forward(x, state: State, rate: int, training: bool, prefix: str = "aspp"):
    b, c, row, col = x.shape
    outs = []
    y = _conv(x, state, prefix + ".branch1.0")
    outs.append(F.relu(_bn(y, state, prefix + ".branch1.1", training, DEFAULT_BN_MOMENTUM)))
    for k, d in ((2, 6 * rate), (3, 12 * rate), (4, 18 * rate)):
        y = _conv(x, state, prefix + ".branch%d.0" % k, 1, d, d)
        outs.append(F.relu(_bn(y, state, prefix + ".branch%d.1" % k, training, DEFAULT_BN_MOMENTUM)))
    g = x.mean(dim=(2, 3), keepdim=True)
    g = _conv(g, state, prefix + ".branch5_conv")
    g = F.relu(_bn(g, state, prefix + ".branch5_bn", training, DEFAULT_BN_MOMENTUM))
    # bilinear(align_corners=True) from a 1x1 map is a broadcast
    outs.append(g.expand(b, g.shape[1], row, col))
    y = _conv(torch.cat(outs, dim=1), state, prefix + ".conv_cat.0")
    return F.relu(_bn(y, state, prefix + ".conv_cat.1", training, DEFAULT_BN_MOMENTUM))


def deeplab_forward(x, state: State, backbone: str = "xception", downsample_factor: int = 16,
                    training: bool = False, dropout: bool = False, return_lowres: bool = False):
    """DeepLab.forward (nets/deeplabv3_plus.py:169-188).  ``dropout`` enables the two
    nn.Dropout layers of cat_conv (only meaningful when training)."""
    H, W = x.shape[2], x.shape[3]
    if backbone == "xception":
        low, hi = xception_forward(x, state, downsample_factor, training)
    elif backbone == "mobilenet":
        low, hi = mobilenet_forward(x, state, downsample_factor, training)
    else:
        raise ValueError("Unsupported backbone - `{}`, Use mobilenet, xception.".format(backbone))
    hi = aspp_forward(hi, state, 16 // downsample_factor, training)
    low = _conv(low, state, "shortcut_conv.0")
    low = F.relu(_bn(low, state, "shortcut_conv.1", training, DEFAULT_BN_MOMENTUM))
    hi = F.interpolate(hi, size=low.shape[2:], mode="bilinear", align_corners=True)
    y = torch.cat((hi, low), dim=1)
    y = _conv(y, state, "cat_conv.0", 1, 1)
    y = F.relu(_bn(y, state, "cat_conv.1", training, DEFAULT_BN_MOMENTUM))
    y = F.dropout(y, 0.5, training and dropout)
    y = _conv(y, state, "cat_conv.4", 1, 1)
    y = F.relu(_bn(y, state, "cat_conv.5", training, DEFAULT_BN_MOMENTUM))
    y = F.dropout(y, 0.1, training and dropout)
    lowres = _conv(y, state, "cls_conv")
    out = F.interpolate(lowres, size=(H, W), mode="bilinear", align_corners=True)
    return (out, lowres) if return_lowres else out


# ----------------------------------------------------------------------------- state builders
def _conv_shapes_xception(num_classes: int, downsample_factor: int):
    """Yield (key, shape, kind) in the reference's state_dict order."""
    items: List[Tuple[str, tuple, str]] = []

    def conv(p, co, ci, k, bias=False, groups=1):
        items.append((p + ".weight", (co, ci // groups, k, k), "conv"))
        if bias:
            items.append((p + ".bias", (co,), "cbias"))

    def bn(p, c):
        items.append((p + ".weight", (c,), "gamma"))
        items.append((p + ".bias", (c,), "beta"))
        items.append((p + ".running_mean", (c,), "rmean"))
        items.append((p + ".running_var", (c,), "rvar"))
        items.append((p + ".num_batches_tracked", (), "nbt"))

    def sep(p, ci, co):
        conv(p + ".depthwise", ci, ci, 3, groups=ci)
        bn(p + ".bn1", ci)
        conv(p + ".pointwise", co, ci, 1)
        bn(p + ".bn2", co)

    def block(p, ci, co, strides, grow_first=True):
        if co != ci or strides != 1:
            conv(p + ".skip", co, ci, 1)
            bn(p + ".skipbn", co)
        f = co if grow_first else ci
        sep(p + ".sepconv1", ci, f)
        sep(p + ".sepconv2", f, co)
        sep(p + ".sepconv3", co, co)

    stride_list = [2, 1, 1] if downsample_factor == 8 else [2, 2, 1]
    b = "backbone"
    conv(b + ".conv1", 32, 3, 3); bn(b + ".bn1", 32)
    conv(b + ".conv2", 64, 32, 3); bn(b + ".bn2", 64)
    block(b + ".block1", 64, 128, 2)
    block(b + ".block2", 128, 256, stride_list[0])
    block(b + ".block3", 256, 728, stride_list[1])
    for i in range(4, 20):
        block(b + ".block%d" % i, 728, 728, 1)
    block(b + ".block20", 728, 1024, stride_list[2], grow_first=False)
    sep(b + ".conv3", 1024, 1536)
    sep(b + ".conv4", 1536, 1536)
    sep(b + ".conv5", 1536, 2048)
    _head_shapes(items, conv, bn, 2048, 256, num_classes)
    return items


def _head_shapes(items, conv, bn, in_ch, low_ch, num_classes):
    a = "aspp"
    conv(a + ".branch1.0", 256, in_ch, 1, bias=True); bn(a + ".branch1.1", 256)
    for k in (2, 3, 4):
        conv(a + ".branch%d.0" % k, 256, in_ch, 3, bias=True); bn(a + ".branch%d.1" % k, 256)
    conv(a + ".branch5_conv", 256, in_ch, 1, bias=True); bn(a + ".branch5_bn", 256)
    conv(a + ".conv_cat.0", 256, 1280, 1, bias=True); bn(a + ".conv_cat.1", 256)
    conv("shortcut_conv.0", 48, low_ch, 1, bias=True); bn("shortcut_conv.1", 48)
    conv("cat_conv.0", 256, 304, 3, bias=True); bn("cat_conv.1", 256)
    conv("cat_conv.4", 256, 256, 3, bias=True); bn("cat_conv.5", 256)
    conv("cls_conv", num_classes, 256, 1, bias=True)


def _conv_shapes_mobilenet(num_classes: int, downsample_factor: int):
    items: List[Tuple[str, tuple, str]] = []

    def conv(p, co, ci, k, bias=False, groups=1):
        items.append((p + ".weight", (co, ci // groups, k, k), "conv"))
        if bias:
            items.append((p + ".bias", (co,), "cbias"))

    def bn(p, c):
        items.append((p + ".weight", (c,), "gamma"))
        items.append((p + ".bias", (c,), "beta"))
        items.append((p + ".running_mean", (c,), "rmean"))
        items.append((p + ".running_var", (c,), "rvar"))
        items.append((p + ".num_batches_tracked", (), "nbt"))

    f = "backbone.features"
    conv(f + ".0.0", 32, 3, 3); bn(f + ".0.1", 32)
    plan = mobilenet_plan(downsample_factor)
    for i in range(1, len(plan)):
        blk = plan[i]
        hidden = round(blk["inp"] * blk["expand"])
        p = f + ".%d" % i
        if blk["expand"] == 1:
            conv(p + ".conv.0", hidden, hidden, 3, groups=hidden); bn(p + ".conv.1", hidden)
            conv(p + ".conv.3", blk["oup"], hidden, 1); bn(p + ".conv.4", blk["oup"])
        else:
            conv(p + ".conv.0", hidden, blk["inp"], 1); bn(p + ".conv.1", hidden)
            conv(p + ".conv.3", hidden, hidden, 3, groups=hidden); bn(p + ".conv.4", hidden)
            conv(p + ".conv.6", blk["oup"], hidden, 1); bn(p + ".conv.7", blk["oup"])
    _head_shapes(items, conv, bn, 320, 24, num_classes)
    return items


def state_schema(backbone: str, num_classes: int = 5, downsample_factor: int = 16):
    if backbone == "xception":
        return _conv_shapes_xception(num_classes, downsample_factor)
    if backbone == "mobilenet":
        return _conv_shapes_mobilenet(num_classes, downsample_factor)
    raise ValueError(backbone)


def make_state(backbone: str, num_classes: int = 5, downsample_factor: int = 16, seed: int = 0,
               randomize_bn_stats: bool = True, conv_std: float | None = None,
               dtype=torch.float32) -> State:
    """Deterministic synthetic weights: conv ~ N(0, std) with
    He-style std = sqrt(2/(k*k*C_out)) so activations stay O(1) through 70+ layers
    (``conv_std`` overrides; the reference's ``weights_init`` uses 0.02), BN gamma ~
    N(1, 0.02), beta ~ N(0, 0.02), conv bias ~ N(0, 0.02); running_mean ~ N(0, 0.1) and
    running_var ~ U(0.5, 1.5) when ``randomize_bn_stats`` so eval-mode BN is not an
    identity.  One CPU generator per tensor keeps any sub-dict reproducible."""
    state: State = {}
    for idx, (key, shape, kind) in enumerate(state_schema(backbone, num_classes, downsample_factor)):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if kind == "conv":
            co, _, k, _ = shape
            std = conv_std if conv_std is not None else math.sqrt(2.0 / (k * k * co))
            t = torch.randn(shape, generator=g, dtype=dtype) * std
        elif kind == "gamma":
            t = 1.0 + 0.02 * torch.randn(shape, generator=g, dtype=dtype)
        elif kind in ("beta", "cbias"):
            t = 0.02 * torch.randn(shape, generator=g, dtype=dtype)
        elif kind == "rmean":
            t = 0.1 * torch.randn(shape, generator=g, dtype=dtype) if randomize_bn_stats \
                else torch.zeros(shape, dtype=dtype)
        elif kind == "rvar":
            t = 0.5 + torch.rand(shape, generator=g, dtype=dtype) if randomize_bn_stats \
                else torch.ones(shape, dtype=dtype)
        elif kind == "nbt":
            t = torch.zeros((), dtype=torch.long)
        else:
            raise AssertionError(kind)
        state[key] = t
    return state


def synthetic_batch(batch: int, size: int = 512, num_classes: int = 5, seed: int = 0,
                    ignore_frac: float = 0.01):
    """Synthetic inputs per SURVEY.md section 8d / dataloader contract (row L):
    images U[0,1) fp32 NCHW, masks int64 in [0,num_classes) with ``ignore_frac`` of the
    pixels set to ``num_classes`` (ignore_index), one-hot labels fp32 [B,H,W,C+1]."""
    g = torch.Generator().manual_seed(1000 + seed)
    imgs = torch.rand(batch, 3, size, size, generator=g)
    pngs = torch.randint(0, num_classes, (batch, size, size), generator=g)
    ign = torch.rand(batch, size, size, generator=g) < ignore_frac
    pngs = torch.where(ign, torch.full_like(pngs, num_classes), pngs)
    labels = torch.eye(num_classes + 1)[pngs]
    return imgs, pngs, labels
