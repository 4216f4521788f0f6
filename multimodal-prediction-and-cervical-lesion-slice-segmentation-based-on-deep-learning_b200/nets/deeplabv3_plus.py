"""DeepLabv3+ on the B200 NHWC engine - drop-in for the reference's ``nets/deeplabv3_plus.py``.

Same public surface as the reference (``DeepLab(num_classes, backbone, pretrained,
downsample_factor)`` :116-188, ``ASPP`` :56-114, ``MobileNetV2`` wrapper :7-49): NCHW float in,
``[B, num_classes, H, W]`` fp32 logits out, identical ``state_dict`` keys/shapes, ``.backbone``
for freezing.  Internally the batch is converted once to NHWC in the engine dtype (bf16 by
default, ``DeepLab.set_compute_dtype(torch.float32)`` for the exact-parity mode) and every
layer runs on the hand-written CUDA kernels behind ``ops.py``; there is no cuDNN/ATen conv.
"""
from functools import partial

import os

import torch
import torch.nn as nn

from .. import ops
from ..backend import get_backend
from .mobilenetv2 import mobilenetv2
from .xception import xception


class MobileNetV2(nn.Module):
    def __init__(self, downsample_factor=8, pretrained=True):
        super().__init__()
        model = mobilenetv2(pretrained)
        self.features = model.features[:-1]
        self.total_idx = len(self.features)
        self.down_idx = [2, 4, 7, 14]
        if downsample_factor == 8:
            for i in range(self.down_idx[-2], self.down_idx[-1]):
                self.features[i].apply(partial(self._nostride_dilate, dilate=2))
            for i in range(self.down_idx[-1], self.total_idx):
                self.features[i].apply(partial(self._nostride_dilate, dilate=4))
        elif downsample_factor == 16:
            for i in range(self.down_idx[-1], self.total_idx):
                self.features[i].apply(partial(self._nostride_dilate, dilate=2))

    @staticmethod
    def _nostride_dilate(m, dilate):
        if isinstance(m, nn.Conv2d):
            if m.stride == (2, 2):
                m.stride = (1, 1)
                if m.kernel_size == (3, 3):
                    m.dilation = (dilate // 2, dilate // 2)
                    m.padding = (dilate // 2, dilate // 2)
            elif m.kernel_size == (3, 3):
                m.dilation = (dilate, dilate)
                m.padding = (dilate, dilate)

    def forward(self, x):
        low_level_features = None
        for i, layer in enumerate(self.features):
            x = layer(x)
            if i == 3:
                low_level_features = x
        return low_level_features, x


_GRAD_CHAIN = os.environ.get("CERVIX_GRAD_CHAIN", "1") != "0"      # A/B switch for tools/profile_step.py


def _conv_bn_relu(x, seq, idx=0, chain=None):
    conv, bn = seq[idx], seq[idx + 1]
    return ops.conv_bn_act(x, conv.weight, conv.stride[0], conv.padding[0], conv.dilation[0], bn, ops.ACT_RELU, None,
                           conv.bias, chain)


class ASPP(nn.Module):
    def __init__(self, dim_in, dim_out, rate=1, bn_mom=0.1):
        super().__init__()

        def branch(k, d):
            return nn.Sequential(nn.Conv2d(dim_in, dim_out, k, 1, padding=0 if k == 1 else d, dilation=d, bias=True),
                                 nn.BatchNorm2d(dim_out, momentum=bn_mom), nn.ReLU(inplace=True))

        self.branch1 = branch(1, rate)
        self.branch2 = branch(3, 6 * rate)
        self.branch3 = branch(3, 12 * rate)
        self.branch4 = branch(3, 18 * rate)
        self.branch5_conv = nn.Conv2d(dim_in, dim_out, 1, 1, 0, bias=True)
        self.branch5_bn = nn.BatchNorm2d(dim_out, momentum=bn_mom)
        self.branch5_relu = nn.ReLU(inplace=True)
        self.conv_cat = nn.Sequential(nn.Conv2d(dim_out * 5, dim_out, 1, 1, padding=0, bias=True),
                                      nn.BatchNorm2d(dim_out, momentum=bn_mom), nn.ReLU(inplace=True))

    def forward(self, x):
        n, row, col, c = x.shape  # NHWC
        # five consumers of x: their gradients are summed inside the data-gradient GEMMs' epilogues (ops.GradChain)
        # instead of by four add passes over the [B,row,col,2048] tensor
        chain = ops.GradChain(5) if (torch.is_grad_enabled() and x.requires_grad and _GRAD_CHAIN) else None
        xs = ops.fanout(x, 5) if chain is not None else (x,) * 5
        outs = [_conv_bn_relu(xi, b, 0, chain) for xi, b in zip(xs, (self.branch1, self.branch2, self.branch3, self.branch4))]
        g = ops.global_avg_pool(xs[4], chain)
        g = ops.conv2d(g, self.branch5_conv.weight, self.branch5_conv.bias, 1, 0, 1)
        g = ops.batchnorm_act(g, self.branch5_bn, ops.ACT_RELU)
        outs.append(ops.broadcast_hw(g, row, col))
        return _conv_bn_relu(ops.cat_channels(outs), self.conv_cat)


class DeepLab(nn.Module):
    def __init__(self, num_classes, backbone="mobilenet", pretrained=True, downsample_factor=16):
        super().__init__()
        if backbone == "xception":
            self.backbone = xception(downsample_factor=downsample_factor, pretrained=pretrained)
            in_channels, low_level_channels = 2048, 256
        elif backbone == "mobilenet":
            self.backbone = MobileNetV2(downsample_factor=downsample_factor, pretrained=pretrained)
            in_channels, low_level_channels = 320, 24
        else:
            raise ValueError('Unsupported backbone - `{}`, Use mobilenet, xception.'.format(backbone))
        self.aspp = ASPP(dim_in=in_channels, dim_out=256, rate=16 // downsample_factor)
        self.shortcut_conv = nn.Sequential(nn.Conv2d(low_level_channels, 48, 1), nn.BatchNorm2d(48),
                                           nn.ReLU(inplace=True))
        self.cat_conv = nn.Sequential(
            nn.Conv2d(48 + 256, 256, 3, stride=1, padding=1), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
            nn.Dropout(0.5),
            nn.Conv2d(256, 256, 3, stride=1, padding=1), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
            nn.Dropout(0.1),
        )
        self.cls_conv = nn.Conv2d(256, num_classes, 1, stride=1)
        self._cervix_dtype = torch.bfloat16

    # -- engine controls (additions; the reference has no equivalent) ----------------------
    def set_compute_dtype(self, dtype: torch.dtype):
        """torch.bfloat16 (tensor-core training path) or torch.float32 (exact-parity path)."""
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("compute dtype must be float32 or bfloat16")
        self._cervix_dtype = dtype
        return self

    def forward_lowres(self, x):
        """Logits before the final x4 bilinear upsample, NHWC in the engine dtype."""
        if x.dtype == torch.uint8:
            # decoded pixels as PIL / numpy hold them, [B,H,W,3] uint8: /255 (utils.preprocess_input) and the layout
            # the stem reads in one pass - the loader's float conversion and NCHW transpose (dataloader.py:40) skipped
            x = get_backend().finish_batch_u8(x.contiguous(), None, 0, self._cervix_dtype)[0]
        else:
            x = ops.to_nhwc(x, self._cervix_dtype)
        low_level_features, x = self.backbone(x)
        x = self.aspp(x)
        low_level_features = _conv_bn_relu(low_level_features, self.shortcut_conv)
        x = ops.upsample_concat(x, low_level_features)       # x4 bilinear written straight into the concat buffer
        x = _conv_bn_relu(x, self.cat_conv, 0)
        x = ops.dropout(x, self.cat_conv[3].p, self.training)
        x = _conv_bn_relu(x, self.cat_conv, 4)
        x = ops.dropout(x, self.cat_conv[7].p, self.training)
        c = self.cls_conv
        return ops.conv2d(x, c.weight, c.bias, 1, 0, 1)

    def forward(self, x):
        H, W = (x.size(1), x.size(2)) if x.dtype == torch.uint8 else (x.size(2), x.size(3))
        return ops.upsample_to_nchw(self.forward_lowres(x), H, W)
