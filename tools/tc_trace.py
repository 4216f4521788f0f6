"""Phase timeline of one CTA-pair tcgen05 launch (first and last cluster): where a short-K pointwise GEMM spends its
cycles between kernel entry, the first operands, each tile's MMAs / epilogue and kernel exit.

    python tools/tc_trace.py [cin cout [stats side]]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from cervix_b200.backend import ConvGeom, get_backend

B = get_backend()
N = 32
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def trace(cin, cout, stats, side, cold=True):
    x = torch.randn(N, 32, 32, cin, device="cuda").bfloat16()
    g = ConvGeom(N, 32, 32, cin, cout, 1, 1, 1, 0, 1)
    wt = torch.randn(cout, cin, 1, 1, device="cuda") * cin ** -0.5
    wp = B.pack_weight(wt, torch.bfloat16, False)
    bias = torch.randn(cout, device="cuda")
    sd = torch.randn(N, 32, 32, cout, device="cuda").bfloat16() if side else None
    ss = torch.ones(cout, device="cuda") if side else None      # the side input comes with a per-channel scale
    for _ in range(3):
        B.conv_fwd_ex(x, wp, bias, g, sd, ss, stats)
    if cold:
        flush.zero_()
    torch.cuda.synchronize()
    B.lib.cvx_debug_tc_trace(1, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    B.conv_fwd_ex(x, wp, bias, g, sd, ss, stats)
    e1.record()
    buf = (C.c_ulonglong * 128)()
    B.lib.cvx_debug_tc_trace(0, buf)
    print("== %d -> %d stats=%s side=%s %s: %.1f us by CUDA events" % (cin, cout, stats, side, "cold L2" if cold else "warm L2", e0.elapsed_time(e1) * 1e3))
    for which, o in (("first cluster", 0), ("last cluster", 64)):
        t = [int(v) for v in buf[o:o + 64]]
        c0 = t[1]
        rel = lambda i: (t[i] - c0) if t[i] else -1   # noqa: E731
        print("  %s: kernel %.1f us by globaltimer, %d cycles" % (which, (t[58] - t[0]) / 1e3, t[57] - c0))
        print("    set-up done %d | dependency wait done %d | first TMA %d | last TMA %d | first operands landed %d" % (rel(2), rel(3), rel(4), rel(5), rel(6)))
        print("    tile:           " + " ".join("%7d" % i for i in range(16) if t[8 + i]))
        print("    MMAs committed  " + " ".join("%7d" % rel(8 + i) for i in range(16) if t[8 + i]))
        print("    accumulator seen" + " ".join("%7d" % rel(24 + i) for i in range(16) if t[24 + i]))
        print("    epilogue done   " + " ".join("%7d" % rel(40 + i) for i in range(16) if t[40 + i]))
        print("    stats flushed %d | exit %d" % (rel(56), rel(57)))


if len(sys.argv) >= 3:
    trace(int(sys.argv[1]), int(sys.argv[2]), len(sys.argv) > 3 and sys.argv[3] == "1", len(sys.argv) > 4 and sys.argv[4] == "1")
else:
    trace(728, 728, False, False)
    trace(728, 728, True, False)
    trace(728, 728, True, False, cold=False)
    trace(728, 728, False, True)
    trace(768, 768, False, False)
    trace(2048, 768, False, False)
