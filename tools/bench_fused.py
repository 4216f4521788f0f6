"""Micro-benchmarks of the fused separable-conv chain kernels on the Xception middle-flow shape (GPU box only).
    python tools/bench_fused.py [--batch 32] [--only name]
Each op is timed with CUDA events over 20 launches after 3 warm-ups, L2 flushed between launches."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from cervix_b200.backend import ConvGeom, get_backend

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--hw", type=int, default=32)
ap.add_argument("--c", type=int, default=728)
ap.add_argument("--only", default="")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--no-flush", action="store_true")
ap.add_argument("--graph", action="store_true", help="time 10 launches replayed as a CUDA graph (no host latency)")
args = ap.parse_args()
B = get_backend()
n, h, c = args.batch, args.hw, args.c
gd = ConvGeom(n, h, h, c, c, 3, 3, 1, 1, 1)
gp = ConvGeom(n, h, h, c, c, 1, 1, 1, 0, 1)
x = torch.randn(n, h, h, c, device="cuda").bfloat16()
dd = torch.randn(n, h, h, c, device="cuda").bfloat16()
add = torch.randn(n, h, h, c, device="cuda").bfloat16()
w9c = torch.randn(9, c, device="cuda")
sc, sh = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
wt = torch.randn(c, c, 1, 1, device="cuda") * (1.0 / c) ** 0.5
wp, wpt = B.pack_weight(wt, torch.bfloat16, False), B.pack_weight(wt, torch.bfloat16, True)
bias = torch.randn(c, device="cuda")
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
tensor_mb = x.numel() * 2 / 1e6
gf = 2.0 * n * h * h * c * c / 1e9

OPS = {
    "dwf_fwd plain": (lambda: B.dwf_fwd(x, w9c, None, None, True, gd, False), 2),
    "dwf_fwd affine+stats": (lambda: B.dwf_fwd(x, w9c, sc, sh, True, gd, True), 2),
    "dw_fwd (old tiled)": (lambda: B.dw_fwd(x, w9c, gd, True), 2),
    "dwf_bwd affine+sums": (lambda: B.dwf_bwd(dd, None, None, None, x, w9c, sc, sh, True, None, gd, True), 3),
    "dwf_bwd affine+sums+side": (lambda: B.dwf_bwd(dd, add, sc, sh, x, w9c, sc, sh, True, None, gd, True), 4),
    "dwf_bwd plain+addend+side": (lambda: B.dwf_bwd(dd, add, sc, sh, x, w9c, None, None, True, add, gd, False), 5),
    "dw_bwd_data (old)": (lambda: B.dw_bwd_data(dd, w9c, x, gd, True), 3),
    "dw_bwd_weight (old)": (lambda: B.dw_bwd_weight(x, dd, gd, True), 2),
    "conv_fwd tc": (lambda: B.conv_fwd(x, wp, bias, gp, True), 2),
    "conv_fwd_ex stats": (lambda: B.conv_fwd_ex(x, wp, bias, gp, None, None, True), 2),
    "conv_dgrad tc": (lambda: B.conv_dgrad(dd, wpt, gp, True), 2),
    "conv_dgrad_ex side": (lambda: B.conv_dgrad_ex(dd, wpt, gp, bias, x, sc), 3),
    "conv_wgrad tc": (lambda: B.conv_wgrad(x, dd, gp, True), 2),
    "bn_stats": (lambda: B.bn_stats(x), 1),
    "affine_act+res": (lambda: B.affine_act(x, sc, sh, add, 1), 3),
    "bn_bwd_sums": (lambda: B.bn_bwd_sums(dd, x, add, 1), 3),
    "bn_bwd_affine": (lambda: B.bn_bwd_affine(dd, None, x, sc, sh, bias, 0, False), 3),
}
print("shape [%d,%d,%d,%d] bf16: %.1f MB per tensor, pointwise GEMM %.1f GF" % (n, h, h, c, tensor_mb, gf))
for name, (fn, passes) in OPS.items():
    if args.only and args.only not in name:
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    if args.graph:
        # GPU time without host launch latency: `inner` back-to-back launches replayed as one CUDA graph
        inner = 10
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            fn()
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(inner):
                    fn()
        torch.cuda.synchronize()
        for _ in range(max(3, args.reps // 4)):
            if not args.no_flush:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / inner)
    else:
        for _ in range(args.reps):
            if not args.no_flush:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    extra = "  %6.0f TF/s" % (gf / ms) if "conv" in name else ""
    print("  %-24s %7.3f ms  %6.0f GB/s (%d passes)%s" % (name, ms, passes * tensor_mb / ms, passes, extra))
