"""GPU tests of the predict / fit entry points: the fused post-processing kernel (csrc/postprocess.cu) against
cv2 and its plain-torch emulation, the ``DeeplabV3`` predictor end to end, and one ``fit_one_epoch`` on the device."""
import os

import numpy as np
import pytest
import torch

from cervix_b200.backend import get_backend
from tests.emu_backend import EmuBackend

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c,h,w,crop,out", [
    (5, 64, 64, (0, 8, 64, 48), (97, 61)),      # up-scaling of a letterboxed crop
    (5, 128, 128, (16, 0, 96, 128), (48, 80)),  # down-scaling
    (21, 32, 48, (0, 0, 32, 48), (32, 48)),     # identity size, many classes
    (2, 17, 19, (3, 2, 11, 13), (1, 1)),        # degenerate output
])
def test_seg_postprocess_matches_cv2(c, h, w, crop, out):
    import cv2
    g = torch.Generator().manual_seed(c * 1000 + h)
    logits = (torch.randn(c, h, w, generator=g) * 3).cuda()
    cls, probs = get_backend().seg_postprocess(logits, crop, out, True)
    ecls, eprobs = EmuBackend().seg_postprocess(logits.cpu(), crop, out, True)
    assert cls.shape == (out[0], out[1]) and cls.dtype == torch.uint8
    assert float((probs.cpu() - eprobs).abs().max()) < 2e-6
    assert (cls.cpu() == ecls).float().mean() >= 0.999
    pr = torch.softmax(logits.cpu().permute(1, 2, 0), -1).numpy()[crop[0]:crop[0] + crop[2], crop[1]:crop[1] + crop[3]]
    ref = cv2.resize(pr, (out[1], out[0]), interpolation=cv2.INTER_LINEAR).reshape(out[0], out[1], c)
    assert np.abs(probs.cpu().numpy() - ref).max() < 1e-5
    assert (cls.cpu().numpy() == ref.argmax(-1)).mean() >= 0.999


@pytest.mark.parametrize("bb", ["xception", "mobilenet"])
def test_predictor_fp32_masks_match_oracle(bb):
    """detect_image / get_miou_png through the fp32 engine vs the CPU oracle network + the reference's cv2 chain."""
    import cv2
    from PIL import Image
    from cervix_b200.deeplab import DeeplabV3
    from cervix_b200.utils.utils import cvtColor, preprocess_input, resize_image
    from oracle import deeplab_ref as O

    state = O.make_state(bb, 5, 16, seed=5)
    pred = DeeplabV3(model_path={k: v.clone() for k, v in state.items()}, backbone=bb, input_shape=[128, 128],
                     cuda=True, compute_dtype=torch.float32)
    rng = np.random.RandomState(1)
    img = Image.fromarray(rng.randint(0, 255, (150, 100, 3), dtype=np.uint8))
    image = cvtColor(img)
    canvas, nw, nh = resize_image(image, (128, 128))
    data = np.expand_dims(np.transpose(preprocess_input(np.array(canvas, np.float32)), (2, 0, 1)), 0)
    with torch.no_grad():
        ref = O.deeplab_forward(torch.from_numpy(data), {k: v.clone() for k, v in state.items()}, bb, 16, False)[0]
    pr = torch.softmax(ref.permute(1, 2, 0), -1).numpy()
    pr = pr[(128 - nh) // 2:(128 - nh) // 2 + nh, (128 - nw) // 2:(128 - nw) // 2 + nw]
    pr = cv2.resize(pr, (100, 150), interpolation=cv2.INTER_LINEAR).argmax(-1)
    got = np.array(pred.get_miou_png(img))
    assert got.shape == (150, 100)
    assert (got == pr).mean() >= 0.999
    assert pred.detect_image(img).size == (100, 150)
    assert 0 < pred.get_FPS(img, 2) < 60


def test_fit_one_epoch_on_device(tmp_path):
    from cervix_b200.nets.deeplabv3_plus import DeepLab
    from cervix_b200.utils.utils_fit import fit_one_epoch

    class Hist:
        def __init__(self):
            self.losses, self.val_loss = [], []

        def append_loss(self, e, a, b):
            self.losses.append(a); self.val_loss.append(b)

    class Ev:
        def on_epoch_end(self, e, m):
            self.seen = e

    torch.manual_seed(0)
    model = DeepLab(5, "xception", False, 16).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    g = torch.Generator().manual_seed(0)

    def batches(n):
        out = []
        for _ in range(n):
            pngs = torch.randint(0, 6, (4, 64, 64), generator=g)
            out.append((torch.rand(4, 3, 64, 64, generator=g), pngs, torch.eye(6)[pngs]))
        return out

    train = batches(1) * 6   # the same batch six times: the loss must go down
    hist, ev = Hist(), Ev()
    first = None
    for epoch in range(2):
        fit_one_epoch(model, model, hist, ev, opt, epoch, 3, 1, train[epoch * 3:epoch * 3 + 3], train[:1], 2, True, True,
                      True, np.array([1, 1, 5, 3, 4], np.float32), 5, True, None, 1, str(tmp_path))
    assert ev.seen == 2 and len(hist.losses) == 2
    assert np.isfinite(hist.losses).all() and hist.losses[1] < hist.losses[0]
    files = os.listdir(tmp_path)
    assert sum(f.startswith("ep00") for f in files) == 2 and "best_epoch_weights.pth" in files
    sd = torch.load(os.path.join(tmp_path, "last_epoch_weights.pth"))
    assert set(sd) == set(model.state_dict())


def test_confusion_matrix_on_device():
    from cervix_b200.utils.utils_metrics import fast_hist, per_class_iu
    g = torch.Generator().manual_seed(0)
    for n, size in ((5, 512 * 512 * 3 + 17), (21, 100003), (2, 1)):
        gt = torch.randint(0, n + 1, (size,), generator=g).to(torch.uint8)
        gt[gt == n] = 255
        pred = torch.randint(0, n, (size,), generator=g).to(torch.uint8)
        ref = EmuBackend().confusion_matrix(pred, gt, n)
        hist = torch.zeros(n, n, dtype=torch.int64, device="cuda")
        fast_hist(gt.cuda(), pred.cuda(), n, hist)
        fast_hist(gt.cuda(), pred.cuda(), n, hist)       # accumulates
        assert (hist.cpu() == 2 * ref).all()
        assert (fast_hist(gt.numpy(), pred.numpy(), n) == ref.numpy()).all()
        assert per_class_iu(hist).shape == (n,)
