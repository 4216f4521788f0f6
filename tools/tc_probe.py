"""Probe the tcgen05 conv kernels case by case, each in its own process (a trapped kernel
poisons the CUDA context), printing error statistics and block-wise error maps that make
descriptor / swizzle mistakes recognisable.  GPU box only:

    python tools/tc_probe.py            # all cases
    python tools/tc_probe.py --one 3    # a single case, in-process
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # kind, n, h, w, cin, cout, k, pad, dil, bias
    ("fwd", 2, 16, 16, 64, 64, 1, 0, 1, False),
    ("fwd", 2, 16, 16, 128, 128, 1, 0, 1, False),
    ("fwd", 2, 32, 32, 728, 728, 1, 0, 1, False),
    ("fwd", 2, 16, 16, 64, 64, 3, 1, 1, False),
    ("fwd", 1, 32, 32, 128, 256, 3, 6, 6, True),
    ("fwd", 2, 24, 40, 304, 256, 3, 1, 1, True),
    ("fwd", 2, 32, 32, 256, 48, 1, 0, 1, True),
    ("fwd", 3, 6, 6, 2048, 256, 3, 12, 12, True),
    ("fwd_s2", 2, 32, 32, 128, 128, 3, 1, 1, False),
    ("fwd_s2", 3, 28, 20, 256, 256, 3, 1, 1, True),
    ("dgrad", 2, 16, 16, 64, 128, 1, 0, 1, False),
    ("dgrad", 2, 32, 32, 728, 728, 1, 0, 1, False),
    ("dgrad", 1, 32, 32, 128, 256, 3, 6, 6, False),
    ("dgrad", 2, 24, 40, 304, 256, 3, 1, 1, False),
    ("wgrad", 2, 16, 16, 64, 64, 1, 0, 1, False),
    ("wgrad", 2, 16, 16, 128, 128, 1, 0, 1, False),
    ("wgrad", 2, 32, 32, 728, 728, 1, 0, 1, False),
    ("wgrad", 2, 16, 16, 64, 128, 3, 1, 1, False),
    ("wgrad", 1, 32, 32, 128, 256, 3, 6, 6, False),
    ("wgrad", 2, 24, 40, 304, 256, 3, 1, 1, False),
    ("wgrad", 2, 32, 32, 256, 48, 1, 0, 1, False),
    ("wgrad", 8, 32, 32, 2048, 256, 3, 12, 12, False),
]


def block_map(err, rb, cb):
    """mean |err| over (rb x cb) blocks of a 2-D tensor, as a small printable grid."""
    import torch
    r, c = err.shape
    r2, c2 = (r // rb) * rb, (c // cb) * cb
    if r2 == 0 or c2 == 0:
        return err.mean().reshape(1, 1)
    return err[:r2, :c2].reshape(r2 // rb, rb, c2 // cb, cb).mean(dim=(1, 3))


def run_one(idx):
    import torch
    from cervix_b200.backend import ConvGeom, get_backend
    from tests.emu_backend import EmuBackend
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    kind, n, h, w, cin, cout, k, pad, dil, bias = CASES[idx]
    B, E = get_backend(), EmuBackend()
    g = ConvGeom(n, h, w, cin, cout, k, k, 2 if kind == "fwd_s2" else 1, pad, dil)
    gen = torch.Generator(device="cuda").manual_seed(idx)
    x = torch.randn((n, h, w, cin), generator=gen, device="cuda").bfloat16()
    wt = torch.randn((cout, cin, k, k), generator=gen, device="cuda") * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn((cout,), generator=gen, device="cuda") if bias else None
    dy = torch.randn((n, g.ho, g.wo, cout), generator=gen, device="cuda").bfloat16()
    if kind in ("fwd", "fwd_s2"):
        wp = B.pack_weight(wt, torch.bfloat16, False)
        got, ref = B.conv_fwd(x, wp, b, g, True), E.conv_fwd(x, wp, b, g, False)
        simt = B.conv_fwd(x, wp, b, g, False)
    elif kind == "dgrad":
        wpt = B.pack_weight(wt, torch.bfloat16, True)
        got, ref = B.conv_dgrad(dy, wpt, g, True), E.conv_dgrad(dy, wpt, g, False)
        simt = B.conv_dgrad(dy, wpt, g, False)
    else:
        got, ref = B.conv_wgrad(x, dy, g, True), E.conv_wgrad(x, dy, g, False)
        simt = B.conv_wgrad(x, dy, g, False)
    torch.cuda.synchronize()
    got, ref, simt = got.float(), ref.float(), simt.float()
    scale = float(ref.abs().max())
    err = (got - ref).abs()
    e = float(err.max()) / scale
    es = float((simt - ref).abs().max()) / scale
    bad = float((err > 2e-2 * scale).float().mean())
    ok = e < (1.5e-2 if kind != "wgrad" else 5e-3)
    print("case %2d %-5s n%d %dx%d cin%d cout%d k%d p%d d%d : rel err %.3e (simt %.3e) bad-frac %.4f %s" %
          (idx, kind, n, h, w, cin, cout, k, pad, dil, e, es, bad, "OK" if ok else "MISMATCH"))
    if not ok:
        e2 = err.reshape(-1, err.shape[-1]) / scale
        torch.set_printoptions(precision=3, linewidth=200, sci_mode=False)
        print("  mean rel err per (32 rows x 16 cols) block, first 8x16 blocks:")
        print(block_map(e2, 32, 16)[:8, :16].cpu())
        print("  mean rel err per row (first 16 rows):", e2[:16].mean(1).cpu())
        print("  mean rel err per col (first 32 cols):", e2.mean(0)[:32].cpu())
        print("  got[0,:8]:", got.reshape(-1, got.shape[-1])[0, :8].cpu())
        print("  ref[0,:8]:", ref.reshape(-1, ref.shape[-1])[0, :8].cpu())
    return ok


def main():
    if "--one" in sys.argv:
        ok = run_one(int(sys.argv[sys.argv.index("--one") + 1]))
        sys.exit(0 if ok else 1)
    fails = 0
    only = None
    if "--kinds" in sys.argv:
        only = sys.argv[sys.argv.index("--kinds") + 1].split(",")
    for i in range(len(CASES)):
        if only and CASES[i][0] not in only:
            continue
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(i)], capture_output=True,
                               text=True, timeout=180)
            out = (p.stdout + p.stderr).strip()
            if p.returncode != 0 and "MISMATCH" not in out:
                out = "case %2d %s CRASHED rc=%d\n%s" % (i, CASES[i], p.returncode, out[-1500:])
        except subprocess.TimeoutExpired:
            out = "case %2d %s TIMEOUT" % (i, CASES[i])
            p = None
        print(out, flush=True)
        if p is None or p.returncode != 0:
            fails += 1
    print("tc_probe: %d / %d cases failed" % (fails, len(CASES)))
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
