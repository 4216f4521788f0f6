"""Aligned Xception backbone on the B200 NHWC engine.

Drop-in for the reference's ``nets/xception.py`` (same constructor, attribute tree and
``state_dict`` keys: SeparableConv2d :9-31, Block :33-73, Xception :76-182), but the forward
pass runs NHWC through the fused CUDA operators in ``ops.py``.  The ``nn.Conv2d`` /
``nn.BatchNorm2d`` children are parameter holders only - their own ``forward`` is never
called - which keeps checkpoints, ``weights_init``, freezing and optimizers working unchanged.

Reference behaviours reproduced on purpose ("reference quirks"):
  * identity-skip blocks add relu(inp) (the in-place relu0 aliases the block input);
  * the low-level feature is the PRE-ReLU output of block2.sepconv2;
  * an unsupported ``downsample_factor`` raises TypeError (the reference formats the ``os``
    module into a '%d'), not ValueError.
"""
import math
import os

import torch
import torch.nn as nn

from .. import ops, ops_fused

bn_mom = 0.0003


def _engine_dtype(module) -> torch.dtype:
    return getattr(module, "_cervix_dtype", torch.bfloat16)


class SeparableConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, dilation=1, bias=False,
                 activate_first=True, inplace=True):
        super().__init__()
        if kernel_size != 3 or bias:
            raise ValueError("cervix_b200 SeparableConv2d supports the reference's 3x3, bias-free form only")
        self.relu0 = nn.ReLU(inplace=inplace)
        self.depthwise = nn.Conv2d(in_channels, in_channels, kernel_size, stride, padding, dilation,
                                   groups=in_channels, bias=bias)
        self.bn1 = nn.BatchNorm2d(in_channels, momentum=bn_mom)
        self.relu1 = nn.ReLU(inplace=True)
        self.pointwise = nn.Conv2d(in_channels, out_channels, 1, 1, 0, 1, 1, bias=bias)
        self.bn2 = nn.BatchNorm2d(out_channels, momentum=bn_mom)
        self.relu2 = nn.ReLU(inplace=True)
        self.activate_first = activate_first

    def forward(self, x, residual=None, out_act=ops.ACT_NONE):
        """x: NHWC.  ``residual``/``out_act`` let the owning Block fuse its skip-add (and the
        next block's leading ReLU) into this layer's last BatchNorm kernel."""
        dw = self.depthwise
        mid_act = ops.ACT_NONE if self.activate_first else ops.ACT_RELU
        y = ops.dwconv3x3(x, dw.weight, dw.stride[0], dw.padding[0], dw.dilation[0], relu_in=self.activate_first)
        y = ops.batchnorm_act(y, self.bn1, mid_act)
        if not self.activate_first:
            out_act = ops.ACT_RELU
        return ops.conv_bn_act(y, self.pointwise.weight, 1, 0, 1, self.bn2, out_act, residual)


# gradient chains for the inputs of the blocks with a 1x1 skip: CERVIX_GRAD_CHAIN=2 (A/B switch)
_GRAD_CHAIN = os.environ.get("CERVIX_GRAD_CHAIN", "1") == "2"


class Block(nn.Module):
    def __init__(self, in_filters, out_filters, strides=1, atrous=None, grow_first=True, activate_first=True,
                 inplace=True):
        super().__init__()
        if atrous is None:
            atrous = [1] * 3
        elif isinstance(atrous, int):
            atrous = [atrous] * 3
        self.head_relu = True
        if out_filters != in_filters or strides != 1:
            self.skip = nn.Conv2d(in_filters, out_filters, 1, stride=strides, bias=False)
            self.skipbn = nn.BatchNorm2d(out_filters, momentum=bn_mom)
            self.head_relu = False
        else:
            self.skip = None
        self.hook_layer = None
        filters = out_filters if grow_first else in_filters
        self.sepconv1 = SeparableConv2d(in_filters, filters, 3, stride=1, padding=1 * atrous[0], dilation=atrous[0],
                                        bias=False, activate_first=activate_first, inplace=self.head_relu)
        self.sepconv2 = SeparableConv2d(filters, out_filters, 3, stride=1, padding=1 * atrous[1], dilation=atrous[1],
                                        bias=False, activate_first=activate_first)
        self.sepconv3 = SeparableConv2d(out_filters, out_filters, 3, stride=strides, padding=1 * atrous[2],
                                        dilation=atrous[2], bias=False, activate_first=activate_first, inplace=inplace)

    def forward(self, inp, inp_is_relu=False, relu_out=False):
        """inp: NHWC.  ``inp_is_relu`` says the producer already applied the ReLU that an
        identity-skip block performs in place on its input; ``relu_out`` asks this block to
        emit relu(out) because its only consumer is such an identity-skip block."""
        fuse = os.environ.get("CERVIX_NO_FUSED_BLOCK") != "1"
        seps = (self.sepconv1, self.sepconv2, self.sepconv3)
        out_act = ops.ACT_RELU if relu_out else ops.ACT_NONE
        if fuse and self.skip is None and inp_is_relu and ops_fused.chain_fusable(seps, inp):
            # training-mode fast path: the whole block as one hand-scheduled chain of fused kernels
            self.hook_layer = None
            return ops_fused.sep_chain(seps, inp, None, True, out_act)
        if self.skip is not None:
            s = self.skip
            fuse3 = fuse and ops_fused.chain_fusable(seps, inp)
            fuse2 = fuse and not fuse3 and ops_fused.chain_fusable(seps[:2], inp)
            # the block input has two consumers (skip conv, separable chain): one shared GradChain lets the later
            # backward add the earlier one's gradient inside its own kernel instead of autograd adding two tensors
            chain = ops.GradChain(2) if ((fuse3 or fuse2) and _GRAD_CHAIN and inp.requires_grad) else None
            a, b = ops.fanout(inp, 2) if chain is not None else (inp, inp)
            if fuse3:
                # stride-1 block with a 1x1 skip (block20): all three fused.  Backward order is forced by the data flow
                # (the chain hands the skip its gradient): the skip conv's data gradient takes the chain's as side input
                skip = ops.conv_bn_act(a, s.weight, s.stride[0], 0, 1, self.skipbn, ops.ACT_NONE, chain=chain)
                self.hook_layer = None
                return ops_fused.sep_chain(seps, b, skip, False, out_act, chain)
            if fuse2:
                # entry-flow blocks: the first two (full-resolution) separable convs fused, their pre-ReLU output
                # materialised once (it is block2's low-level feature), the strided third one on the operator path.
                # The skip conv is created AFTER the chain so that its backward runs first (autograd runs ready nodes
                # latest-created first) and the chain's depthwise backward adds its gradient as the addend.
                x = ops_fused.sep_chain(seps[:2], b, None, False, ops.ACT_NONE, chain)
                skip = ops.conv_bn_act(a, s.weight, s.stride[0], 0, 1, self.skipbn, ops.ACT_NONE, chain=chain)
                self.hook_layer = x
                return self.sepconv3(x, residual=skip, out_act=out_act)
            skip = ops.conv_bn_act(inp, s.weight, s.stride[0], 0, 1, self.skipbn, ops.ACT_NONE)
        else:
            if not inp_is_relu:
                inp = ops.relu(inp)
            skip = inp
        x = self.sepconv1(inp)
        x = self.sepconv2(x)
        self.hook_layer = x
        return self.sepconv3(x, residual=skip, out_act=ops.ACT_RELU if relu_out else ops.ACT_NONE)


class Xception(nn.Module):
    def __init__(self, downsample_factor):
        super().__init__()
        if downsample_factor == 8:
            stride_list = [2, 1, 1]
        elif downsample_factor == 16:
            stride_list = [2, 2, 1]
        else:
            raise ValueError('xception.py: output stride=%d is not supported.' % os)  # TypeError, as upstream
        self.conv1 = nn.Conv2d(3, 32, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(32, momentum=bn_mom)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(32, 64, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(64, momentum=bn_mom)

        self.block1 = Block(64, 128, 2)
        self.block2 = Block(128, 256, stride_list[0], inplace=False)
        self.block3 = Block(256, 728, stride_list[1])
        rate = 16 // downsample_factor
        for i in range(4, 20):
            setattr(self, "block%d" % i, Block(728, 728, 1, atrous=rate))
        self.block20 = Block(728, 1024, stride_list[2], atrous=rate, grow_first=False)
        self.conv3 = SeparableConv2d(1024, 1536, 3, 1, 1 * rate, dilation=rate, activate_first=False)
        self.conv4 = SeparableConv2d(1536, 1536, 3, 1, 1 * rate, dilation=rate, activate_first=False)
        self.conv5 = SeparableConv2d(1536, 2048, 3, 1, 1 * rate, dilation=rate, activate_first=False)
        self.layers = []

        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def forward(self, x):
        """x: NHWC engine tensor (DeepLab.forward converts the NCHW image batch)."""
        self.layers = []
        x = ops.conv2d(x, self.conv1.weight, None, 2, 1, 1)
        x = ops.batchnorm_act(x, self.bn1, ops.ACT_RELU)
        x = ops.conv_bn_act(x, self.conv2.weight, 1, 1, 1, self.bn2, ops.ACT_RELU)
        x = self.block1(x)
        x = self.block2(x)
        low_level = ops.cut_point(self.block2.hook_layer, "low_level")
        x = self.block3(x, relu_out=True)            # block4 is an identity-skip block
        # end of the entry flow: 3 % of the parameters are behind this point, ~35 % of the backward time (the 128^2 - 256^2
        # tensors) - the data-parallel trainer all-reduces the other 97 % of the gradient under that part of backward
        x = ops.cut_point(x, "entry_out")
        for i in range(4, 20):
            x = getattr(self, "block%d" % i)(x, inp_is_relu=True, relu_out=(i < 19))
        x = self.block20(x)
        x = self.conv3(x)
        x = self.conv4(x)
        x = self.conv5(x)
        return low_level, x


def xception(pretrained=True, downsample_factor=16):
    model = Xception(downsample_factor=downsample_factor)
    if pretrained:
        path = os.path.join("model_data", "xception_pytorch_imagenet.pth")
        if not os.path.exists(path):
            raise FileNotFoundError(
                "pretrained=True needs %s (the reference downloads it from GitHub; this build has no network)" % path)
        model.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
    return model
