"""Micro-benchmarks of the bandwidth kernels on the layer shapes of the Xception step (GPU box):
achieved GB/s against the algorithmic bytes.  python tools/bench_kernels.py [--batch 32]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cervix_b200.backend import ConvGeom, get_backend

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=32); args = ap.parse_args()
B = get_backend(); N = args.batch
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

print("depthwise 3x3 (bf16): shape, fwd / bwd-data / bwd-weight  ms (GB/s of in+out bytes)")
for (c, hw, s) in [(64, 256, 1), (128, 256, 1), (128, 256, 2), (256, 128, 1), (728, 64, 1), (728, 32, 1), (1536, 32, 1)]:
    g = ConvGeom(N, hw, hw, c, c, 3, 3, s, 1, 1)
    x = torch.randn((N, hw, hw, c), device="cuda").bfloat16()
    dy = torch.randn((N, g.ho, g.wo, c), device="cuda").bfloat16()
    w = torch.randn((9, c), device="cuda")
    byts = (x.numel() + dy.numel()) * 2
    t0 = timeit(lambda: B.dw_fwd(x, w, g, True)); t1 = timeit(lambda: B.dw_bwd_data(dy, w, x, g, True))
    t2 = timeit(lambda: B.dw_bwd_weight(x, dy, g, True))
    print("  C=%4d %3dx%-3d s%d : %.3f (%.0f)  %.3f (%.0f)  %.3f (%.0f)" %
          (c, hw, hw, s, t0, byts / t0 / 1e6, t1, (byts + x.numel() * 2) / t1 / 1e6, t2, byts / t2 / 1e6))

print("batchnorm (bf16, train, relu): fwd / bwd ms (GB/s: fwd 3 passes, bwd 7 passes of the tensor)")
for (c, hw) in [(64, 256), (128, 256), (256, 128), (728, 32), (2048, 32), (256, 128)]:
    x = torch.randn((N, hw, hw, c), device="cuda").bfloat16()
    gam = torch.ones(c, device="cuda"); bet = torch.zeros(c, device="cuda")
    rm = torch.zeros(c, device="cuda"); rv = torch.ones(c, device="cuda")
    y, m, i = B.bn_forward(x, None, gam, bet, rm, rv, 1, True, 0.1, 1e-5)
    t0 = timeit(lambda: B.bn_forward(x, None, gam, bet, rm, rv, 1, True, 0.1, 1e-5))
    t1 = timeit(lambda: B.bn_backward(x, x, y, gam, m, i, 1, True, False))
    nb = x.numel() * 2
    print("  C=%4d %3dx%-3d : %.3f (%.0f)  %.3f (%.0f)" % (c, hw, hw, t0, 3 * nb / t0 / 1e6, t1, 7 * nb / t1 / 1e6))

print("tcgen05 conv (bf16): fwd / dgrad / wgrad ms (TFLOP/s)")
for (cin, cout, hw, k, d) in [(728, 728, 32, 1, 1), (128, 128, 256, 1, 1), (256, 256, 128, 1, 1), (1536, 2048, 32, 1, 1),
                              (2048, 256, 32, 3, 12), (304, 256, 128, 3, 1), (256, 256, 128, 3, 1), (64, 128, 256, 1, 1)]:
    g = ConvGeom(N, hw, hw, cin, cout, k, k, 1, d * (k // 2), d)
    x = torch.randn((N, hw, hw, cin), device="cuda").bfloat16()
    dy = torch.randn((N, hw, hw, cout), device="cuda").bfloat16()
    wp = torch.randn((k * k, cout, cin), device="cuda").bfloat16(); wpt = torch.randn((k * k, cin, cout), device="cuda").bfloat16()
    fl = 2.0 * N * hw * hw * cin * cout * k * k
    t0 = timeit(lambda: B.conv_fwd(x, wp, None, g, True)); t1 = timeit(lambda: B.conv_dgrad(dy, wpt, g, True))
    t2 = timeit(lambda: B.conv_wgrad(x, dy, g, True))
    print("  %4d->%4d %3dx%-3d k%d d%-2d : %.3f (%.0f)  %.3f (%.0f)  %.3f (%.0f)" %
          (cin, cout, hw, hw, k, d, t0, fl / t0 / 1e9, t1, fl / t1 / 1e9, t2, fl / t2 / 1e9))
