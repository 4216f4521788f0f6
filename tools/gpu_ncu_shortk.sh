#!/bin/bash
# short-K pointwise conv (entry flow 64 -> 128 @256^2, K = one 64-channel block): event timing of the plain / stats
# variants, then one ncu --set full capture each with source correlation
set -u
mkdir -p gpurun_out
python - > gpurun_out/shortk_plain.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from cervix_b200.backend import ConvGeom, get_backend
B = get_backend()
def t(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (n, h, cin, cout) in ((32, 256, 64, 128), (32, 256, 128, 128), (32, 128, 128, 256), (32, 128, 256, 256), (768, 64, 64, 256), (768, 64, 256, 64)):
    g = ConvGeom(n, h, h, cin, cout, 1, 1, 1, 0, 1)
    x = torch.randn((n, h, h, cin), device="cuda").bfloat16()
    w = torch.randn(cout, cin, 1, 1, device="cuda") * 0.05
    wp = B.pack_weight(w, torch.bfloat16, False)
    bias = torch.randn(cout, device="cuda")
    mb = (x.numel() + n * h * h * cout) * 2 / 1e6
    a = t(lambda: B.conv_fwd(x, wp, None, g, True))
    b = t(lambda: B.conv_fwd_ex(x, wp, None, g, None, None, True))
    c = t(lambda: B.conv_fwd_act(x, wp, bias, g, 1))
    print("%dx%dx%dx%d -> %d: plain %.0f us (%.0f GB/s)  stats %.0f us  bias+relu %.0f us   [%.0f MB]" % (n, h, h, cin, cout, a, mb / a * 1e3, b, c, mb))
PY
cat gpurun_out/shortk_plain.log
python tools/run_conv_once.py 32 fwd 64 128 1 256 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_fwd -s 2 -c 1 -o gpurun_out/ncu_shortk_plain \
    python tools/run_conv_once.py 32 fwd 64 128 1 256 > gpurun_out/ncu_shortk_plain.log 2>&1
echo "rc=$?"
python tools/run_conv_once.py 32 fwd_ex 64 128 1 256 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_fwd -s 2 -c 1 -o gpurun_out/ncu_shortk_stats \
    python tools/run_conv_once.py 32 fwd_ex 64 128 1 256 > gpurun_out/ncu_shortk_stats.log 2>&1
echo "rc=$?"
