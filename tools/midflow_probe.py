"""Where does the middle-flow pointwise GEMM (M = 32*32*32 pixels, N = K = 728) lose its time?  Times the three passes
for varied K, N, channel alignment and epilogue variant through the public backend calls (L2 flushed between launches).

    python tools/midflow_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from cervix_b200.backend import ConvGeom, get_backend

B = get_backend()
torch.manual_seed(0)
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()   # noqa: E731


def timed(fn, reps=7):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


N = 32
print("%-34s %9s %9s %9s %9s %9s   TF/s fwd" % ("cin -> cout @32x32 x32", "fwd", "fwd+stat", "dgrad", "dgrad+sd", "wgrad"))
for cin, cout in ((728, 728), (768, 768), (384, 728), (1456, 728), (728, 256), (728, 512), (728, 1024), (2048, 768), (768, 2048)):
    x, dd = bf(N, 32, 32, cin), bf(N, 32, 32, cout)
    sidex = bf(N, 32, 32, cin)
    g = ConvGeom(N, 32, 32, cin, cout, 1, 1, 1, 0, 1)
    wt = torch.randn(cout, cin, 1, 1, device=dev) * cin ** -0.5
    wp, wpt = B.pack_weight(wt, torch.bfloat16, False), B.pack_weight(wt, torch.bfloat16, True)
    bias_o, bias_i, sc_i = torch.randn(cout, device=dev), torch.randn(cin, device=dev), torch.rand(cin, device=dev) + 0.5
    t_f = timed(lambda: B.conv_fwd_ex(x, wp, bias_o, g, None, None, False))
    t_fs = timed(lambda: B.conv_fwd_ex(x, wp, bias_o, g, None, None, True))
    t_d = timed(lambda: B.conv_dgrad_ex(dd, wpt, g, None, None, None))
    t_ds = timed(lambda: B.conv_dgrad_ex(dd, wpt, g, bias_i, sidex, sc_i))
    t_w = timed(lambda: B.conv_wgrad(x, dd, g, True))
    gf = 2.0 * N * 32 * 32 * cin * cout
    print("%-34s %8.1fu %8.1fu %8.1fu %8.1fu %8.1fu   %6.0f %6.0f %6.0f %6.0f %6.0f" % (
        "%d -> %d" % (cin, cout), t_f, t_fs, t_d, t_ds, t_w, *(gf / t / 1e6 for t in (t_f, t_fs, t_d, t_ds, t_w))), flush=True)
    del x, dd, sidex
