"""One representative launch of every kernel family of the training step at its batch-32 shape (ncu target for the round-2
evidence; also prints CUDA-event times, L2 flushed between launches, with the algorithmic bytes / flops of each launch).

    python tools/run_families_once.py [--only name]

Algorithmic bytes = every tensor the kernel must touch, once; flops = 2*M*N*K."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from cervix_b200.backend import ConvGeom, get_backend

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
B = get_backend()
torch.manual_seed(0)
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()   # noqa: E731


def timed(name, fn, nbytes=None, flops=None):
    if args.only and args.only not in name:
        return
    for _ in range(2):
        fn()
    ts = []
    for _ in range(args.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    extra = ""
    if nbytes:
        extra += "  %7.1f MB algorithmic  %6.0f GB/s" % (nbytes / 1e6, nbytes / ms / 1e6)
    if flops:
        extra += "  %7.1f GF  %6.0f TFLOP/s" % (flops / 1e9, flops / ms / 1e9)
    print("%-52s %8.1f us%s" % (name, ms * 1e3, extra), flush=True)


N = 32
# ---- middle-flow tensor [32,32,32,728]
c = 728
x, dd, add = bf(N, 32, 32, c), bf(N, 32, 32, c), bf(N, 32, 32, c)
w9c = torch.randn(9, c, device=dev)
sc, sh, bias = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.1, torch.randn(c, device=dev)
gd = ConvGeom(N, 32, 32, c, c, 3, 3, 1, 1, 1)
gp = ConvGeom(N, 32, 32, c, c, 1, 1, 1, 0, 1)
wt = torch.randn(c, c, 1, 1, device=dev) * c ** -0.5
wp, wpt = B.pack_weight(wt, torch.bfloat16, False), B.pack_weight(wt, torch.bfloat16, True)
T = x.numel() * 2
gf = 2.0 * N * 32 * 32 * c * c
timed("dwf_fwd affine+stats (mid flow)", lambda: B.dwf_fwd(x, w9c, sc, sh, True, gd, True), 2 * T)
timed("dwf_bwd affine+sums (mid flow)", lambda: B.dwf_bwd(dd, None, None, None, x, w9c, sc, sh, True, None, gd, True), 3 * T)
timed("dwf_bwd plain+addend (mid flow, block input)", lambda: B.dwf_bwd(dd, None, None, None, x, w9c, None, None, True, add, gd, False), 4 * T)
timed("conv_fwd_ex stats 728->728 (mid.pw fwd)", lambda: B.conv_fwd_ex(x, wp, None, gp, None, None, True), 2 * T, gf)
timed("conv_dgrad_ex side 728->728 (mid.pw dgrad)", lambda: B.conv_dgrad_ex(dd, wpt, gp, bias, x, sc), 3 * T, gf)
timed("conv_wgrad tc 728x728 (mid.pw wgrad)", lambda: B.conv_wgrad(x, dd, gp, True), 2 * T, gf)
timed("bn_bwd_sums (colreduce BnRawBwdF)", lambda: B.bn_bwd_sums(dd, x, add, 1), 3 * T)
timed("bn_bwd_affine", lambda: B.bn_bwd_affine(dd, None, x, sc, sh, bias, 0, False), 3 * T)
timed("affine_act + residual", lambda: B.affine_act(x, sc, sh, add, 1), 3 * T)
timed("bn_stats (colreduce StatsF)", lambda: B.bn_stats(x), T)
del x, dd, add
# ---- decoder tensor [32,128,128,256]: plain conv + BatchNorm layers
c = 256
x, dy = bf(N, 128, 128, c), bf(N, 128, 128, c)
gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.1
rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
T = x.numel() * 2
timed("bn_forward train+relu (StatsF + bn_apply)", lambda: B.bn_forward(x, None, gamma, beta, rm, rv, 1, True, 0.1, 1e-5), 3 * T)
mean, invstd = torch.zeros(c, device=dev), torch.ones(c, device=dev)
timed("bn_backward mask-from-x (BnBwdF + bn_bwd_apply)", lambda: B.bn_backward(dy, x, None, gamma, mean, invstd, 1, True, False, beta), 5 * T)
g3 = ConvGeom(N, 128, 128, 304, 256, 3, 3, 1, 1, 1)
x3, wp3 = bf(N, 128, 128, 304), bf(9, 256, 304)
timed("conv_fwd tc 3x3 304->256 @128 (cat_conv.0)", lambda: B.conv_fwd(x3, wp3, None, g3, True), x3.numel() * 2 + T,
      2.0 * N * 128 * 128 * 256 * 304 * 9)
del x3, wp3
timed("upsample_fwd x4 (32^2 -> 128^2, 256 ch)", lambda: B.upsample_fwd(x[:, :32, :32].contiguous(), 128, 128), T + T / 16)
low48 = bf(N, 128, 128, 48)
xs32 = x[:, :32, :32].contiguous()
timed("upsample_concat (x4 into the 304-channel concat buffer)", lambda: B.upsample_concat(xs32, low48), T + T / 16 + 2 * low48.numel() * 2)
step_dev = torch.ones(1, dtype=torch.int32, device=dev)
timed("dropout_fwd, no stored mask (decoder, 256 ch @128^2)", lambda: B.dropout_fwd(x, 0.5, 1234, step_dev, want_mask=False), 2 * T)
timed("dropout_bwd_seeded", lambda: B.dropout_bwd_seeded(dy, 0.5, 1234, step_dev), 2 * T)
g2048 = bf(N, 1, 1, 2048)
timed("spatial_broadcast (ASPP pooled branch backward, 2048 ch @32^2)", lambda: B.spatial_broadcast(g2048, 32, 32, 1.0 / 1024), N * 32 * 32 * 2048 * 2)
del x, dy, low48, xs32
# ---- training augmentation: one batch of 32 VOC-sized decoded images onto 512 x 512 canvases (four launches)
import numpy as np  # noqa: E402
from cervix_b200.utils import dataloader as D  # noqa: E402
_rng = np.random.RandomState(0)
_items = []
np.random.seed(0)
for _i in range(N):
    _ih, _iw = (375, 500) if _i % 2 == 0 else (500, 375)
    _items.append((_rng.randint(0, 256, (_ih, _iw, 3)).astype(np.uint8), _rng.randint(0, 6, (_ih, _iw)).astype(np.uint8),
                   D.draw_params(_iw, _ih, (512, 512))))
_plan = D.pack_batch(_items, (512, 512))
_aug = D.DeviceAugmenter(dev)
_blobs = _aug.upload(_plan, non_blocking=False)
timed("augment batch (resize_rows + compose + blur5 + rotate_jitter)", lambda: _aug.run(_plan, _blobs),
      int(_plan.src.numel()) + N * 512 * 512 * 4)
# ---- strided / dilated depthwise (entry flow stride 2, exit flow dilation 2)
g2 = ConvGeom(N, 256, 256, 128, 128, 3, 3, 2, 1, 1)
x2, w2 = bf(N, 256, 256, 128), torch.randn(9, 128, device=dev)
timed("dw_fwd stride 2 (block1, 128 ch @256^2)", lambda: B.dw_fwd(x2, w2, g2, True), x2.numel() * 2 * 1.25)
del x2
g4 = ConvGeom(N, 32, 32, 1536, 1536, 3, 3, 1, 2, 2)
x4, w4 = bf(N, 32, 32, 1536), torch.randn(9, 1536, device=dev)
timed("dw_fwd dilation 2 (exit flow, 1536 ch @32^2)", lambda: B.dw_fwd(x4, w4, g4, True), x4.numel() * 4)
del x4
# ---- loss: logits fp32 NCHW [32,5,512,512], int64 class map
logits = torch.randn(N, 5, 512, 512, device=dev)
tgt = torch.randint(0, 6, (N, 512, 512), device=dev)
cw = torch.tensor([1., 1, 5, 3, 4], device=dev)
stats = B.seg_loss_stats(logits, tgt, None, cw, 0.5, 2.0, 0.5)
gvec = torch.tensor([0., 1., 1., 0.], device=dev)
timed("seg_loss_stats (CE + focal + dice + f_score sums)", lambda: B.seg_loss_stats(logits, tgt, None, cw, 0.5, 2.0, 0.5),
      logits.numel() * 4 + tgt.numel() * 8)
timed("seg_loss_grad", lambda: B.seg_loss_grad(logits, tgt, None, cw, stats, gvec, 0.5, 2.0, 1.0, 1e-5),
      logits.numel() * 8 + tgt.numel() * 8)
lo = bf(N, 128, 128, 5)
timed("upsample_to_nchw_fwd (logits x4 -> fp32 NCHW)", lambda: B.upsample_to_nchw_fwd(lo, 512, 512), logits.numel() * 4 + lo.numel() * 2)
timed("upsample_to_nchw_bwd", lambda: B.upsample_to_nchw_bwd(logits, 128, 128, torch.bfloat16), logits.numel() * 4 + lo.numel() * 2)
del logits, tgt
# ---- optimizer on the 54.7 M flat parameters
n = 54_709_448
p, g, m, v = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
v.abs_()
hyper = torch.tensor([1e-4, 0.9, 0.999, 1e-8, 0.0, 1.0], device=dev)
step = torch.ones(1, dtype=torch.int32, device=dev)
timed("adam_dev (54.7 M parameters)", lambda: B.adam_step_dev(p, g, m, v, hyper, step), n * 28)
del p, g, m, v
# ---- classifier: grouped GEMM of the SAGE layer (16 patients, 4 modalities) and the encoder's dominant conv variants
xs = [torch.randn(256 if i < 3 else 64, 1024, device=dev) for i in range(8)]
ws = [torch.randn(512, 1024, device=dev) * 0.03 for _ in range(8)]
ys = [torch.empty(t.shape[0], 512, device=dev) for t in xs]
probs = [dict(a=a, b=w, c=y, m=a.shape[0], n=512, k=1024, lda_m=1024, lda_k=1, ldb_n=1024, ldb_k=1, ldc=512)
         for a, w, y in zip(xs, ws, ys)]
timed("gemm_grouped SAGE layer (8 problems, K=1024)", lambda: B.gemm_grouped(probs), None,
      sum(2.0 * q["m"] * 512 * 1024 for q in probs))
ge = ConvGeom(256, 64, 64, 64, 256, 1, 1, 1, 0, 1)       # ResNet-101 layer1 1x1 64 -> 256 on 256 patches
xe, we = bf(256, 64, 64, 64), bf(1, 256, 64)
be = torch.zeros(256, device=dev)
timed("encoder conv_fwd_act 1x1 64->256 @64^2 x256 (folded BN + relu)", lambda: B.conv_fwd_act(xe, we, be, ge, 1),
      xe.numel() * 2 * 5, 2.0 * 256 * 64 * 64 * 64 * 256)
ge3 = ConvGeom(256, 16, 16, 256, 256, 3, 3, 1, 1, 1)      # layer3 3x3 256 -> 256 (23 blocks)
xe3, we3 = bf(256, 16, 16, 256), bf(9, 256, 256)
timed("encoder conv_fwd_act 3x3 256->256 @16^2 x256", lambda: B.conv_fwd_act(xe3, we3, be, ge3, 1),
      xe3.numel() * 4, 2.0 * 256 * 16 * 16 * 256 * 256 * 9)
