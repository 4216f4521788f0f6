"""ResNet-101 patch encoder of the classifier's image branches on the B200 NHWC engine.

Reference: ``MultiModal Prediction/Graph_Structure(data_augmentation).py`` :136-168 -
``models.resnet101(pretrained=True)`` with ``fc = Linear(2048, 1024)``, ``eval()``, applied to 16
patches (256x256, cut x-major from the image resized to 1024x1024) of each modality, one patch per
call.  Here the same network (identical ``state_dict`` keys/shapes as torchvision's ``resnet101``, so
ImageNet checkpoints load unchanged) runs all patches of all modalities as ONE batch through the
tcgen05 implicit-GEMM convolutions; BatchNorm uses its running statistics (the reference never
trains or un-freezes this network).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..backend import ConvGeom, get_backend


def _fold_bn(conv_weight: torch.Tensor, bn: nn.BatchNorm2d):
    """Eval-mode BatchNorm folded into the preceding bias-free conv: w' = w * g/sqrt(var+eps), b' = beta - mean*g/sqrt(..)."""
    scale = bn.weight.detach().float() * torch.rsqrt(bn.running_var.float() + bn.eps)
    w = conv_weight.detach().float() * scale.view(-1, 1, 1, 1)
    b = bn.bias.detach().float() - bn.running_mean.float() * scale
    return w, b.contiguous()


class _FoldedConv:
    """Packed bf16 weights + fp32 bias of one conv+bn pair of the frozen encoder, rebuilt when its tensors change."""

    def __init__(self, conv: nn.Conv2d, bn: nn.BatchNorm2d):
        self.conv, self.bn = conv, bn
        self.key = None
        self.wp = self.bias = None

    def get(self):
        t = (self.conv.weight, self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var)
        key = tuple((x.data_ptr(), x._version) for x in t)
        if key != self.key:
            w, self.bias = _fold_bn(self.conv.weight, self.bn)
            self.wp = get_backend().pack_weight(w.contiguous(), torch.bfloat16, False)
            self.key = key
        return self.wp, self.bias


class _FoldedStem(_FoldedConv):
    """The 7x7/2 stem of the frozen encoder as an im2col + 1x1 tensor-core conv with bn1 folded in: the filter becomes a
    [cout, 152] matrix over the K-padded patch vector (tap-major, 3 channels per tap), bias + ReLU in the epilogue."""

    def get(self):
        t = (self.conv.weight, self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var)
        key = tuple((x.data_ptr(), x._version) for x in t)
        if key != self.key:
            w, self.bias = _fold_bn(self.conv.weight, self.bn)
            cout, cin, kh, kw = w.shape
            k = kh * kw * cin
            self.kpad = (k + 7) // 8 * 8
            w2 = torch.nn.functional.pad(w.permute(0, 2, 3, 1).reshape(cout, k), (0, self.kpad - k))
            self.wp = get_backend().pack_weight(w2.reshape(cout, self.kpad, 1, 1).contiguous(), torch.bfloat16, False)
            self.key = key
        return self.wp, self.bias

    def __call__(self, x):
        B = get_backend()
        wp, bias = self.get()
        conv = self.conv
        cout, cin, kh, kw = conv.weight.shape
        n, h, w, _ = x.shape
        g = ConvGeom(n, h, w, cin, cout, kh, kw, conv.stride[0], conv.padding[0], 1)
        patches = B.im2col_narrow(x.contiguous(), g, self.kpad)
        g1 = ConvGeom(n, g.ho, g.wo, self.kpad, cout, 1, 1, 1, 0, 1)
        return B.conv_fwd_act(patches, wp, bias, g1, ops.ACT_RELU, None, None)


def _conv_act(x, fc: _FoldedConv, act: int, residual=None, ones=None):
    """act(conv(x) + folded bn (+ residual)) as ONE tensor-core kernel (cvx_conv_fwd_tc_act)."""
    B = get_backend()
    conv = fc.conv
    wp, bias = fc.get()
    cout, cin, kh, kw = conv.weight.shape
    stride, pad = conv.stride[0], conv.padding[0]
    if stride != 1 and kh == 1:        # 1x1 stride-s conv == 1x1 conv of the subsampled input
        x = B.subsample(x, stride)
        stride = 1
    n, h, w, _ = x.shape
    g = ConvGeom(n, h, w, cin, cout, kh, kw, stride, pad, 1)
    return B.conv_fwd_act(x.contiguous(), wp, bias, g, act, residual, ones)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)   # torchvision v1.5: stride on the 3x3
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        self._folded = None

    def forward_folded(self, x, ones):
        """Inference fast path (eval mode, bf16 engine, no autograd): three fused conv+bn(+residual)+relu kernels."""
        if self._folded is None:
            self._folded = [_FoldedConv(self.conv1, self.bn1), _FoldedConv(self.conv2, self.bn2),
                            _FoldedConv(self.conv3, self.bn3),
                            _FoldedConv(self.downsample[0], self.downsample[1]) if self.downsample is not None else None]
        f1, f2, f3, fd = self._folded
        identity = x if fd is None else _conv_act(x, fd, ops.ACT_NONE)
        y = _conv_act(x, f1, ops.ACT_RELU)
        y = _conv_act(y, f2, ops.ACT_RELU)
        return _conv_act(y, f3, ops.ACT_RELU, identity, None)

    def forward(self, x):
        identity = x
        if self.downsample is not None:
            d = self.downsample[0]
            identity = ops.conv2d(x, d.weight, None, d.stride[0], 0, 1)
            identity = ops.batchnorm_act(identity, self.downsample[1], ops.ACT_NONE)
        y = ops.conv2d(x, self.conv1.weight, None, 1, 0, 1)
        y = ops.batchnorm_act(y, self.bn1, ops.ACT_RELU)
        y = ops.conv2d(y, self.conv2.weight, None, self.conv2.stride[0], 1, 1)
        y = ops.batchnorm_act(y, self.bn2, ops.ACT_RELU)
        y = ops.conv2d(y, self.conv3.weight, None, 1, 0, 1)
        return ops.batchnorm_act(y, self.bn3, ops.ACT_RELU, identity)   # relu(bn3(.) + identity)


class ResNet101Encoder(nn.Module):
    """torchvision ``resnet101`` with ``fc -> Linear(2048, out_features)``; NCHW fp32 in, [N, out] fp32 out."""

    def __init__(self, out_features: int = 1024, layers=(3, 4, 23, 3)):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._make_layer(64, layers[0], 1)
        self.layer2 = self._make_layer(128, layers[1], 2)
        self.layer3 = self._make_layer(256, layers[2], 2)
        self.layer4 = self._make_layer(512, layers[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(2048, out_features)
        self._cervix_dtype = torch.bfloat16
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def _make_layer(self, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes * 4:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, 1, stride, bias=False),
                                       nn.BatchNorm2d(planes * 4))
        layers = [Bottleneck(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * 4
        layers += [Bottleneck(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def set_compute_dtype(self, dtype: torch.dtype):
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("compute dtype must be float32 or bfloat16")
        self._cervix_dtype = dtype
        return self

    def _fold_ok(self, x) -> bool:
        B = get_backend()
        return (not self.training and not torch.is_grad_enabled() and self._cervix_dtype == torch.bfloat16 and x.is_cuda
                and getattr(B, "is_sm100", lambda: False)() and hasattr(B, "conv_fwd_act"))

    def forward(self, x):
        return self.forward_nhwc(ops.to_nhwc(x, self._cervix_dtype))

    def encode_images(self, images, new_size: int = 1024, patch_size: int = 256):
        """[B,3,H,W] fp32 images in [0,1], or the decoded uint8 images [B,H,W,3] (then the resize is Pillow's own 8-bit
        BILINEAR resampler, bit for bit what the reference's ``img.resize`` produces) -> [B*16, out] features: resize +
        x-major patch split + ImageNet normalisation as one kernel straight into the stem's NHWC layout
        (``split_patches_nhwc``), then the encoder."""
        return self.forward_nhwc(split_patches_nhwc(images, new_size, patch_size, self._cervix_dtype))

    def forward_nhwc(self, x):
        """Encoder body on an NHWC tensor already in the compute dtype."""
        fold = self._fold_ok(x)
        c1 = self.conv1
        if fold and c1.weight.shape[1] <= 4:
            if getattr(self, "_stem", None) is None:
                self._stem = _FoldedStem(self.conv1, self.bn1)
            x = self._stem(x)                     # conv1 + bn1 + relu: im2col, then one tensor-core kernel
        else:
            x = ops.conv2d_narrow_in(x, c1.weight, c1.stride[0], c1.padding[0])
            x = ops.batchnorm_act(x, self.bn1, ops.ACT_RELU)
        x = ops.maxpool3x3s2(x)
        if fold:
            # frozen eval-mode network (the only way the reference uses it): BatchNorm folded into the conv weights,
            # bias + residual + ReLU in the GEMM epilogue - no separate normalisation pass over any activation
            if getattr(self, "_ones", None) is None or self._ones.device != x.device:
                self._ones = torch.ones(2048, dtype=torch.float32, device=x.device)
            for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
                for blk in layer:
                    x = blk.forward_folded(x, self._ones)
        else:
            for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
                for blk in layer:
                    x = blk(x)
        x = ops.global_avg_pool(x)                                        # [N,1,1,2048]
        w = self.fc.weight.reshape(self.fc.out_features, self.fc.in_features, 1, 1)
        y = ops.conv2d(x.float() if self._cervix_dtype == torch.float32 else x, w, self.fc.bias, 1, 0, 1)
        return y.reshape(y.shape[0], -1).float()


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def split_patches_nhwc(images: torch.Tensor, new_size: int = 1024, patch_size: int = 256,
                       dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """Tensor form of ``resize_and_split_image`` (:151-161) + ToTensor/Normalize (:145-148) as ONE kernel
    (``cvx_split_patches``): [B,3,H,W] fp32 in [0,1] -> [B*16, 256, 256, 3] NHWC in the encoder's compute dtype,
    patches enumerated x-major (outer loop over columns) as the reference's list comprehension does."""
    if images.dtype == torch.uint8:
        # the decoded image itself, [B,H,W,3] as PIL / numpy hold it: Pillow's own fixed-point resampler, bit for bit
        return get_backend().split_patches_u8(images.contiguous(), new_size, patch_size, IMAGENET_MEAN, IMAGENET_STD, dtype)
    return get_backend().split_patches(images.contiguous(), new_size, patch_size, IMAGENET_MEAN, IMAGENET_STD, dtype)


def split_patches(images: torch.Tensor, new_size: int = 1024, patch_size: int = 256) -> torch.Tensor:
    """NCHW fp32 view of ``split_patches_nhwc`` ([B*16, 3, 256, 256]) for callers that want the layout the
    reference's ``transform(patch)`` produces; the encoder itself is fed by ``ResNet101Encoder.encode_images``."""
    return split_patches_nhwc(images, new_size, patch_size, torch.float32).permute(0, 3, 1, 2).contiguous()


def resize_and_split_image(image, target_size: int = 1024, split_size: int = 256):
    """PIL form of the reference helper (:151-161): path or PIL image -> bilinear resize to target_size^2 -> the
    16 PIL patches in the reference's order (outer loop over x)."""
    from PIL import Image
    img = Image.open(image) if isinstance(image, (str, bytes)) or hasattr(image, "__fspath__") else image
    resized = img.resize((target_size, target_size), Image.BILINEAR)
    return [resized.crop((i, j, i + split_size, j + split_size))
            for i in range(0, target_size, split_size) for j in range(0, target_size, split_size)]


def _pil_to_tensor(image) -> torch.Tensor:
    import numpy as np
    a = np.asarray(image.convert("RGB"), dtype=np.float32) / 255.0
    return torch.from_numpy(a).permute(2, 0, 1).contiguous()


@torch.no_grad()
def extract_features(image, model: ResNet101Encoder, transform=None, device=None):
    """Two call forms.
      * reference form (:164-168): ``extract_features(pil_patch, model, transform, device) -> np.ndarray[1024]``
        (``transform`` defaults to ToTensor + ImageNet normalisation);
      * batched form: ``extract_features(patches[P,3,256,256], model) -> Tensor[P,1024]`` - all patches of all
        modalities in ONE pass through the encoder, which is how the B200 path is meant to be driven."""
    model.eval()
    if torch.is_tensor(image):
        return model(image)
    dev = device if device is not None else next(model.parameters()).device
    if transform is not None:
        x = transform(image)
    else:
        mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
        std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
        x = (_pil_to_tensor(image) - mean) / std
    return model(x.unsqueeze(0).to(dev)).float().cpu().numpy().flatten()


@torch.no_grad()
def extract_patient_features(image, model: ResNet101Encoder, device=None):
    """All 16 patch features of one colposcopic image in one encoder pass: PIL image / path -> np.ndarray[16, 1024]
    (what the reference's per-patch loop, :196-200, accumulates into ``patch_features_dict``)."""
    model.eval()
    dev = device if device is not None else next(model.parameters()).device
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    x = torch.stack([_pil_to_tensor(p) for p in resize_and_split_image(image)])
    return model(((x - mean) / std).to(dev)).float().cpu().numpy()
