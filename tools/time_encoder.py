"""Wall / device time of the frozen patch encoder over 3 x 16 images (768 patches), four repetitions (GPU box only)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from cervix_b200.multimodal.patch_encoder import ResNet101Encoder
torch.manual_seed(0)
enc = ResNet101Encoder().cuda().eval()
imgs = torch.rand(3, 16, 3, 512, 512, device="cuda")
with torch.no_grad():
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(3):
            f = enc.encode_images(imgs[i])
        e1.record(); torch.cuda.synchronize()
        print("rep %d: %.2f ms device, %.2f ms wall" % (rep, e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
