"""Per-shape timing of every tensor-core convolution of DeepLabv3+ Xception ds=16 at the bench
batch: forward, data gradient and weight gradient through the C ABI (GPU box only).
    python tools/bench_conv_shapes.py [--batch 32] [--reps 10] [--no-flush] [--only substr]
Each launch is timed with CUDA events (L2 flushed between launches unless --no-flush); the last column is the
per-step cost = time x count of that shape in the network."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from cervix_b200.backend import ConvGeom, get_backend

# (name, cin, cout, k, stride, dil, h_in, count)
SHAPES = [
    ("conv2 3x3 32->64 @256", 32, 64, 3, 1, 1, 256, 1),
    ("b1.skip 1x1s2 64->128 @256", 64, 128, 1, 2, 1, 256, 1),
    ("b1.pw 64->128 @256", 64, 128, 1, 1, 1, 256, 1),
    ("b1.pw 128->128 @256", 128, 128, 1, 1, 1, 256, 1),
    ("b1.pw 128->128 @128", 128, 128, 1, 1, 1, 128, 1),
    ("b2.skip 1x1s2 128->256 @128", 128, 256, 1, 2, 1, 128, 1),
    ("b2.pw 128->256 @128", 128, 256, 1, 1, 1, 128, 1),
    ("b2.pw 256->256 @128", 256, 256, 1, 1, 1, 128, 1),
    ("b2.pw 256->256 @64", 256, 256, 1, 1, 1, 64, 1),
    ("b3.skip 1x1s2 256->728 @64", 256, 728, 1, 2, 1, 64, 1),
    ("b3.pw 256->728 @64", 256, 728, 1, 1, 1, 64, 1),
    ("b3.pw 728->728 @64", 728, 728, 1, 1, 1, 64, 1),
    ("mid.pw 728->728 @32", 728, 728, 1, 1, 1, 32, 50),
    ("b20 728->1024 @32", 728, 1024, 1, 1, 1, 32, 2),
    ("b20.pw 1024->1024 @32", 1024, 1024, 1, 1, 1, 32, 1),
    ("conv3.pw 1024->1536 @32", 1024, 1536, 1, 1, 1, 32, 1),
    ("conv4.pw 1536->1536 @32", 1536, 1536, 1, 1, 1, 32, 1),
    ("conv5.pw 1536->2048 @32", 1536, 2048, 1, 1, 1, 32, 1),
    ("aspp.b1 2048->256 @32", 2048, 256, 1, 1, 1, 32, 1),
    ("aspp.b2 3x3d6 2048->256", 2048, 256, 3, 1, 6, 32, 1),
    ("aspp.b3 3x3d12 2048->256", 2048, 256, 3, 1, 12, 32, 1),
    ("aspp.b4 3x3d18 2048->256", 2048, 256, 3, 1, 18, 32, 1),
    ("aspp.cat 1280->256 @32", 1280, 256, 1, 1, 1, 32, 1),
    ("shortcut 256->48 @128", 256, 48, 1, 1, 1, 128, 1),
    ("cat_conv.0 3x3 304->256 @128", 304, 256, 3, 1, 1, 128, 1),
    ("cat_conv.4 3x3 256->256 @128", 256, 256, 3, 1, 1, 128, 1),
]

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--only", default="")
ap.add_argument("--no-flush", action="store_true")
args = ap.parse_args()
B = get_backend()
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(args.reps):
        if not args.no_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / args.reps


n = args.batch
tot_ms = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
tot_gf = 0.0
print("%-32s %8s | %8s %6s | %8s %6s | %8s %6s | %8s" % ("shape", "GF", "fwd ms", "TF/s", "dgrad ms", "TF/s",
                                                        "wgrad ms", "TF/s", "step ms"))
for name, cin, cout, k, s, dil, h, count in SHAPES:
    if args.only and args.only not in name:
        continue
    pad = dil * (k // 2)
    g = ConvGeom(n, h, h, cin, cout, k, k, s, pad, dil)
    x = torch.randn(n, h, h, cin, device="cuda").bfloat16()
    dy = torch.randn(n, g.ho, g.wo, cout, device="cuda").bfloat16()
    w = torch.randn(cout, cin, k, k, device="cuda") * (1.0 / (cin * k * k)) ** 0.5
    wp, wpt = B.pack_weight(w, torch.bfloat16, False), B.pack_weight(w, torch.bfloat16, True)
    gf = 2.0 * n * g.ho * g.wo * cout * cin * k * k / 1e9
    t_f = timeit(lambda: B.conv_fwd(x, wp, None, g, True))
    if s == 1:
        t_d = timeit(lambda: B.conv_dgrad(dy, wpt, g, True))
        t_w = timeit(lambda: B.conv_wgrad(x, dy, g, True))
    else:  # stride-2 skips run dgrad/wgrad on the subsampled tensor as stride-1 1x1 convs
        g1 = ConvGeom(n, g.ho, g.wo, cin, cout, 1, 1, 1, 0, 1)
        xs = x[:, ::2, ::2].contiguous()
        t_d = timeit(lambda: B.conv_dgrad(dy, wpt, g1, True))
        t_w = timeit(lambda: B.conv_wgrad(xs, dy, g1, True))
    step = (t_f + t_d + t_w) * count
    tot_ms["fwd"] += t_f * count
    tot_ms["dgrad"] += t_d * count
    tot_ms["wgrad"] += t_w * count
    tot_gf += 3 * gf * count
    print("%-32s %8.1f | %8.3f %6.0f | %8.3f %6.0f | %8.3f %6.0f | %8.2f" % (
        name, gf, t_f, gf / t_f, t_d, gf / t_d, t_w, gf / t_w, step), flush=True)
    del x, dy, w, wp, wpt
tsum = sum(tot_ms.values())
print("total per step: fwd %.2f ms, dgrad %.2f ms, wgrad %.2f ms = %.2f ms for %.1f TF -> %.0f TF/s" % (
    tot_ms["fwd"], tot_ms["dgrad"], tot_ms["wgrad"], tsum, tot_gf / 1e3, tot_gf / tsum))
