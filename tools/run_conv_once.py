"""Run the decoder 3x3 conv (304->256 @128x128) a few times through the C ABI - the target of
the ncu captures (tensor-core forward kernel).  python tools/run_conv_once.py [batch] [kind]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cervix_b200.backend import ConvGeom, get_backend
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
kind = sys.argv[2] if len(sys.argv) > 2 else "fwd"
B = get_backend()
g = ConvGeom(N, 128, 128, 304, 256, 3, 3, 1, 1, 1)
x = torch.randn((N, 128, 128, 304), device="cuda").bfloat16()
dy = torch.randn((N, 128, 128, 256), device="cuda").bfloat16()
wp = torch.randn((9, 256, 304), device="cuda").bfloat16()
for _ in range(3):
    if kind == "fwd":
        B.conv_fwd(x, wp, None, g, True)
    elif kind == "wgrad":
        B.conv_wgrad(x, dy, g, True)
torch.cuda.synchronize()
print("ok")
