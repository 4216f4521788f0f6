"""Run one dense conv shape a few times through the C ABI - the target of the ncu captures.
    python tools/run_conv_once.py [batch] [fwd|dgrad|wgrad|fwd_ex] [cin cout k h]
default shape = decoder 3x3 conv (304->256 @128x128)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cervix_b200.backend import ConvGeom, get_backend
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
kind = sys.argv[2] if len(sys.argv) > 2 else "fwd"
cin, cout, k, h = (int(v) for v in sys.argv[3:7]) if len(sys.argv) > 6 else (304, 256, 3, 128)
B = get_backend()
g = ConvGeom(N, h, h, cin, cout, k, k, 1, k // 2, 1)
x = torch.randn((N, h, h, cin), device="cuda").bfloat16()
dy = torch.randn((N, h, h, cout), device="cuda").bfloat16()
w = torch.randn(cout, cin, k, k, device="cuda") * 0.02
wp, wpt = B.pack_weight(w, torch.bfloat16, False), B.pack_weight(w, torch.bfloat16, True)
for _ in range(3):
    if kind == "fwd":
        B.conv_fwd(x, wp, None, g, True)
    elif kind == "fwd_ex":
        B.conv_fwd_ex(x, wp, None, g, None, None, True)
    elif kind == "dgrad":
        B.conv_dgrad(dy, wpt, g, True)
    elif kind == "wgrad":
        B.conv_wgrad(x, dy, g, True)
torch.cuda.synchronize()
print("ok")
