"""Launch the round-1 v8 kernels a few times at their batch-32 shapes (ncu target; also prints CUDA-event times).
    python tools/run_new_kernels_once.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from cervix_b200.backend import ConvGeom, get_backend

B = get_backend()
torch.manual_seed(0)
dev = "cuda"


def timed(name, fn, nbytes, reps=5):
    for _ in range(2):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print("%-44s %8.1f us  %7.0f MB algorithmic  %6.0f GB/s" % (name, ms * 1e3, nbytes / 1e6, nbytes / ms / 1e6))


# stride-2 depthwise data gradient, entry flow block1 (128 ch, 256^2 -> 128^2)
g = ConvGeom(32, 256, 256, 128, 128, 3, 3, 2, 1, 1)
x = torch.randn(32, 256, 256, 128, device=dev).bfloat16()
dy = torch.randn(32, 128, 128, 128, device=dev).bfloat16()
w9c = torch.randn(9, 128, device=dev)
timed("dw_s2_dgrad 32x256x256x128 (relu mask)", lambda: B.dw_bwd_data(dy, w9c, x, g, True),
      (x.numel() * 2 + dy.numel()) * 2)
del x, dy
# stem weight gradient (3 -> 32, 3x3 stride 2 at 512^2)
g = ConvGeom(32, 512, 512, 3, 32, 3, 3, 2, 1, 1)
x = torch.randn(32, 512, 512, 3, device=dev).bfloat16()
dy = torch.randn(32, 256, 256, 32, device=dev).bfloat16()
timed("stem_wgrad_mma 32x512x512x3 -> 32", lambda: B.conv_wgrad(x, dy, g, False), (x.numel() + dy.numel()) * 2)
del x, dy
# final upsample gradient (fp32 NCHW 5x512x512 -> NHWC bf16 128x128x5)
dy = torch.randn(32, 5, 512, 512, device=dev)
timed("upsample_to_nchw_bwd 32x5x512x512", lambda: B.upsample_to_nchw_bwd(dy, 128, 128, torch.bfloat16),
      dy.numel() * 4 + 32 * 128 * 128 * 5 * 2)
del dy
# classifier patch pipeline: 48 images 512^2 -> 768 patches 256^2 NHWC bf16
imgs = torch.rand(48, 3, 512, 512, device=dev)
timed("split_patches 48x3x512x512 -> 768x256x256x3", lambda: B.split_patches(imgs, 1024, 256, (0.485, 0.456, 0.406),
                                                                            (0.229, 0.224, 0.225), torch.bfloat16),
      imgs.numel() * 4 + 768 * 256 * 256 * 3 * 2)
del imgs
# BatchNorm backward with the mask recomputed from x (decoder 256 ch at 128^2)
c = 256
x = torch.randn(32, 128, 128, c, device=dev).bfloat16()
dy = torch.randn(32, 128, 128, c, device=dev).bfloat16()
gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.1
mean, invstd = torch.zeros(c, device=dev), torch.ones(c, device=dev)
timed("bn_backward mask-from-x 32x128x128x256", lambda: B.bn_backward(dy, x, None, gamma, mean, invstd, 1, True, False, beta),
      x.numel() * 2 * 5)
