"""``DeeplabV3`` predictor - drop-in for the reference's ``deeplab.py`` (class :58, detect_image :108-209,
get_FPS :211-240, get_miou_png :304-345): PIL image in, PIL image out, same letterbox + ``/255`` pre-processing and
the same softmax -> crop -> bilinear resize -> argmax order, so masks agree with the reference pixel for pixel
(up to fp32 rounding at ties).

Differences underneath: the network is the B200 NHWC engine (``nets.deeplabv3_plus.DeepLab``), and the whole
post-processing chain is ONE CUDA kernel on the logits (``cvx_seg_postprocess``); only the uint8 class map is copied
to the host, not the [H,W,C] fp32 probability tensor the reference hands to ``cv2.resize``.  ONNX export
(deeplab.py:245-302) is outside the hot path and not provided.
"""
import colorsys
import copy
import time

import numpy as np
import torch
from PIL import Image

from .backend import get_backend
from .nets.deeplabv3_plus import DeepLab
from .utils.utils import cvtColor, letterbox_geometry, preprocess_input, resize_image, show_config

_VOC_COLORS = [(0, 0, 0), (128, 128, 0), (128, 0, 0), (0, 128, 0), (0, 0, 128), (128, 0, 128), (0, 128, 128),
               (128, 128, 128), (64, 0, 0), (192, 0, 0), (64, 128, 0), (192, 128, 0), (64, 0, 128), (192, 0, 128),
               (64, 128, 128), (192, 128, 128), (0, 64, 0), (128, 64, 0), (0, 192, 0), (128, 192, 0), (0, 64, 128),
               (128, 64, 12)]


class DeeplabV3(object):
    _defaults = {
        "model_path": "logs/best_epoch_weights.pth",
        "num_classes": 5,
        "backbone": "xception",
        "input_shape": [512, 512],
        "downsample_factor": 16,
        "mix_type": 1,          # 0: blend with the input, 1: colour mask only, 2: keep the foreground pixels
        "cuda": True,
        "compute_dtype": None,  # extra: torch.float32 for the exact-parity engine (default bf16 on the GPU)
    }

    @classmethod
    def get_defaults(cls, n):
        return cls._defaults.get(n, "Unrecognized attribute name '" + n + "'")

    def __init__(self, **kwargs):
        self.__dict__.update(self._defaults)
        for name, value in kwargs.items():
            setattr(self, name, value)
        if self.num_classes <= 21:
            self.colors = list(_VOC_COLORS)
        else:
            hsv = [(x / self.num_classes, 1.0, 1.0) for x in range(self.num_classes)]
            self.colors = [tuple(int(v * 255) for v in colorsys.hsv_to_rgb(*t)) for t in hsv]
        self.generate()
        show_config(**{k: getattr(self, k) for k in self._defaults})

    # ------------------------------------------------------------------ model
    def generate(self, onnx=False):
        if onnx:
            raise NotImplementedError("ONNX export is outside the B200 hot path (reference deeplab.py:245-302)")
        self.net = DeepLab(num_classes=self.num_classes, backbone=self.backbone,
                           downsample_factor=self.downsample_factor, pretrained=False)
        state = self.model_path
        if not isinstance(state, dict):
            state = torch.load(state, map_location="cpu")
        self.net.load_state_dict(state)
        self.net = self.net.eval()
        if self.compute_dtype is not None:
            self.net.set_compute_dtype(self.compute_dtype)
        if self.cuda:
            self.net = self.net.cuda()
        print("{} model, and classes loaded.".format("state_dict" if isinstance(self.model_path, dict) else self.model_path))

    # ------------------------------------------------------------------ shared pipeline
    def _prepare(self, image):
        image = cvtColor(image)
        w, h = self.input_shape[1], self.input_shape[0]
        canvas, nw, nh = resize_image(image, (w, h))
        data = np.expand_dims(np.transpose(preprocess_input(np.array(canvas, np.float32)), (2, 0, 1)), 0)
        images = torch.from_numpy(data)
        if self.cuda:
            images = images.cuda()
        return image, images, nw, nh

    def _class_map(self, images, nw, nh, out_hw):
        """uint8 class map [out_h, out_w] on the host: network + fused softmax/crop/resize/argmax."""
        with torch.no_grad():
            logits = self.net(images)[0].float().contiguous()
            top, left = (self.input_shape[0] - nh) // 2, (self.input_shape[1] - nw) // 2
            cls, _ = get_backend().seg_postprocess(logits, (top, left, nh, nw), out_hw)
            return cls.cpu().numpy()

    # ------------------------------------------------------------------ public API
    def detect_image(self, image, count=False, name_classes=None):
        image, images, nw, nh = self._prepare(image)
        old_img = copy.deepcopy(image)
        oh, ow = np.array(image).shape[0], np.array(image).shape[1]
        pr = self._class_map(images, nw, nh, (oh, ow))
        if count:
            classes_nums = np.zeros([self.num_classes])
            total = oh * ow
            print("-" * 63)
            print("|%25s | %15s | %15s|" % ("Key", "Value", "Ratio"))
            print("-" * 63)
            for i in range(self.num_classes):
                num = int(np.sum(pr == i))
                if num > 0:
                    print("|%25s | %15s | %14.2f%%|" % (str(name_classes[i]), str(num), num / total * 100))
                    print("-" * 63)
                classes_nums[i] = num
            print("classes_nums:", classes_nums)
        if self.mix_type == 0:
            seg = np.array(self.colors, np.uint8)[pr.reshape(-1)].reshape(oh, ow, -1)
            return Image.blend(old_img, Image.fromarray(seg), 0.7)
        if self.mix_type == 1:
            seg = np.array(self.colors, np.uint8)[pr.reshape(-1)].reshape(oh, ow, -1)
            return Image.fromarray(seg)
        if self.mix_type == 2:
            seg = (np.expand_dims(pr != 0, -1) * np.array(old_img, np.float32)).astype("uint8")
            return Image.fromarray(seg)
        return image

    def get_FPS(self, image, test_interval):
        """Seconds per image of network + post-processing (the reference times softmax + argmax + crop)."""
        _, images, nw, nh = self._prepare(image)
        self._class_map(images, nw, nh, (nh, nw))
        if self.cuda:
            torch.cuda.synchronize()
        t1 = time.time()
        for _ in range(test_interval):
            self._class_map(images, nw, nh, (nh, nw))   # the D->H copy of the class map synchronises
        t2 = time.time()
        return (t2 - t1) / test_interval

    def get_miou_png(self, image):
        image, images, nw, nh = self._prepare(image)
        oh, ow = np.array(image).shape[0], np.array(image).shape[1]
        return Image.fromarray(np.uint8(self._class_map(images, nw, nh, (oh, ow))))
