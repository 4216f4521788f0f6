"""MobileNetV2 backbone on the B200 NHWC engine.

Drop-in for the reference's ``nets/mobilenetv2.py`` (conv_bn :10-15, conv_1x1_bn :17-22,
InvertedResidual :24-72, MobileNetV2 :74-135 with the same ``features.N.conv.M`` key schema).
``nn.Conv2d`` / ``nn.BatchNorm2d`` children only hold parameters and geometry (the
stride->dilation rewrite of deeplabv3_plus.py:18-43 mutates ``m.stride/dilation/padding`` on
them and is honoured here); compute goes through ``ops.py``.
"""
import math
import os

import torch
import torch.nn as nn

from .. import ops

BatchNorm2d = nn.BatchNorm2d


def _conv_bn_act(x, conv: nn.Conv2d, bn: nn.BatchNorm2d, act, residual=None):
    if conv.groups == 1:
        y = ops.conv2d(x, conv.weight, conv.bias, conv.stride[0], conv.padding[0], conv.dilation[0])
    else:
        y = ops.dwconv3x3(x, conv.weight, conv.stride[0], conv.padding[0], conv.dilation[0], relu_in=False)
    return ops.batchnorm_act(y, bn, act, residual)


class _ConvBNReLU6(nn.Sequential):
    """[Conv2d, BatchNorm2d, ReLU6] with the reference's sequential indices 0/1/2."""

    def forward(self, x):
        return _conv_bn_act(x, self[0], self[1], ops.ACT_RELU6)


def conv_bn(inp, oup, stride):
    return _ConvBNReLU6(nn.Conv2d(inp, oup, 3, stride, 1, bias=False), BatchNorm2d(oup), nn.ReLU6(inplace=True))


def conv_1x1_bn(inp, oup):
    return _ConvBNReLU6(nn.Conv2d(inp, oup, 1, 1, 0, bias=False), BatchNorm2d(oup), nn.ReLU6(inplace=True))


class InvertedResidual(nn.Module):
    def __init__(self, inp, oup, stride, expand_ratio):
        super().__init__()
        self.stride = stride
        assert stride in [1, 2]
        hidden_dim = round(inp * expand_ratio)
        self.use_res_connect = self.stride == 1 and inp == oup
        layers = []
        if expand_ratio != 1:
            layers += [nn.Conv2d(inp, hidden_dim, 1, 1, 0, bias=False), BatchNorm2d(hidden_dim), nn.ReLU6(inplace=True)]
        layers += [nn.Conv2d(hidden_dim, hidden_dim, 3, stride, 1, groups=hidden_dim, bias=False),
                   BatchNorm2d(hidden_dim), nn.ReLU6(inplace=True),
                   nn.Conv2d(hidden_dim, oup, 1, 1, 0, bias=False), BatchNorm2d(oup)]
        self.conv = nn.Sequential(*layers)

    def forward(self, x):
        seq = self.conv
        y = x
        i = 0
        while i < len(seq):
            conv, bn = seq[i], seq[i + 1]
            last = i + 2 >= len(seq)
            if last:  # linear bottleneck: BN only, residual add fused into it
                y = _conv_bn_act(y, conv, bn, ops.ACT_NONE, x if self.use_res_connect else None)
                i += 2
            else:
                y = _conv_bn_act(y, conv, bn, ops.ACT_RELU6)
                i += 3
        return y


class MobileNetV2(nn.Module):
    def __init__(self, n_class=1000, input_size=224, width_mult=1.):
        super().__init__()
        input_channel = 32
        last_channel = 1280
        setting = [  # t, c, n, s
            [1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2], [6, 320, 1, 1],
        ]
        assert input_size % 32 == 0
        input_channel = int(input_channel * width_mult)
        self.last_channel = int(last_channel * width_mult) if width_mult > 1.0 else last_channel
        feats = [conv_bn(3, input_channel, 2)]
        for t, c, n, s in setting:
            output_channel = int(c * width_mult)
            for i in range(n):
                feats.append(InvertedResidual(input_channel, output_channel, s if i == 0 else 1, expand_ratio=t))
                input_channel = output_channel
        feats.append(conv_1x1_bn(input_channel, self.last_channel))
        self.features = nn.Sequential(*feats)
        self.classifier = nn.Sequential(nn.Dropout(0.2), nn.Linear(self.last_channel, n_class))
        self._initialize_weights()

    def forward(self, x):
        raise NotImplementedError("the ImageNet classifier head is outside the DeepLab hot path; "
                                  "use nets.deeplabv3_plus.MobileNetV2 (features only)")

    def _initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()


def mobilenetv2(pretrained=False, **kwargs):
    model = MobileNetV2(n_class=1000, **kwargs)
    if pretrained:
        path = os.path.join("model_data", "mobilenet_v2.pth.tar")
        if not os.path.exists(path):
            raise FileNotFoundError(
                "pretrained=True needs %s (the reference downloads it from GitHub; this build has no network)" % path)
        model.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
    return model
