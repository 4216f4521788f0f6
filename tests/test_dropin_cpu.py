"""The reference scripts' import surface resolved through ``<pkg>/dropin``: module paths,
``fit_one_epoch`` (signature, checkpoint files, loss bookkeeping) and the ``DeeplabV3`` predictor, driven on the CPU by
the plain-torch emulation of the C ABI.  The predictor's mask is compared with the reference's own post-processing
order (softmax -> crop -> cv2.resize -> argmax, deeplab.py:141-154) applied to the same logits."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import cervix_b200.backend as backend
from tests.emu_backend import EmuBackend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "multimodal-prediction-and-cervical-lesion-slice-segmentation-based-on-deep-learning_b200", "dropin")


@pytest.fixture(autouse=True)
def emu():
    prev = backend.set_backend(EmuBackend())
    yield
    backend.set_backend(prev)


def test_reference_import_paths_resolve():
    """Only <pkg>/dropin on sys.path, cwd elsewhere - exactly what the reference's scripts would see."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from nets.deeplabv3_plus import DeepLab\n"
        "from nets.deeplabv3_training import CE_Loss, Dice_loss, Focal_Loss, weights_init, get_lr_scheduler, set_optimizer_lr\n"
        "from utils.utils_fit import fit_one_epoch\n"
        "from utils.utils_metrics import f_score\n"
        "from utils.utils import cvtColor, preprocess_input, resize_image, show_config, get_lr, seed_everything\n"
        "from deeplab import DeeplabV3\n"
        "from my_mae_model import fusion_model_mae_2\n"
        "from mae_utils import generate_mask\n"
        "from util import Logger, adjust_learning_rate\n"
        "import inspect\n"
        "print(list(inspect.signature(fit_one_epoch).parameters))\n" % DROPIN)
    out = subprocess.run([sys.executable, "-c", code], cwd="/tmp", capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip().splitlines()[-1] == str(
        ["model_train", "model", "loss_history", "eval_callback", "optimizer", "epoch", "epoch_step", "epoch_step_val",
         "gen", "gen_val", "Epoch", "cuda", "dice_loss", "focal_loss", "cls_weights", "num_classes", "fp16", "scaler",
         "save_period", "save_dir", "local_rank"])


def test_two_and_three_modal_script_imports_resolve():
    """``from my_mae_model_2_NL import fusion_model_mae_2`` etc. (Two_Modal/train(*).py:30, Three_Modal/train(*).py:30):
    each module name gives the class with that file's defaults and its 156-entry state_dict."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import inspect\n"
        "for name, t in (('my_mae_model_2', 2), ('my_mae_model_2_NL', 2), ('my_mae_model_2_AL', 2), ('my_mae_model_2_NA', 2),\n"
        "                ('my_mae_model_three', 3), ('my_mae_model', 4)):\n"
        "    mod = __import__(name)\n"
        "    m = mod.fusion_model_mae_2(1024, 512, 512, 0.3)\n"
        "    print(name, m.train_type_num, len(m.state_dict()), m._DEFAULT_MIX)\n"
        "    assert m.train_type_num == t\n" % DROPIN)
    out = subprocess.run([sys.executable, "-c", code], cwd="/tmp", capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0] == "my_mae_model_2 2 156 False" and lines[4] == "my_mae_model_three 3 156 False"
    assert lines[5] == "my_mae_model 4 148 True"


class _History:
    def __init__(self):
        self.losses, self.val_loss = [], []

    def append_loss(self, epoch, loss, val_loss):
        self.losses.append(loss)
        self.val_loss.append(val_loss)


class _Eval:
    def __init__(self):
        self.calls = []

    def on_epoch_end(self, epoch, model):
        self.calls.append(epoch)


def _batches(n, size, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        imgs = torch.rand(2, 3, size, size, generator=g)
        pngs = torch.randint(0, 6, (2, size, size), generator=g)
        out.append((imgs, pngs, torch.eye(6)[pngs]))
    return out


def test_fit_one_epoch_matches_hand_rolled_loop(tmp_path):
    from cervix_b200.nets.deeplabv3_plus import DeepLab
    from cervix_b200.nets.deeplabv3_training import Dice_loss, Focal_Loss
    from cervix_b200.utils.utils_fit import fit_one_epoch

    cls_w = np.array([1, 1, 5, 3, 4], np.float32)
    train, val = _batches(2, 32, 0), _batches(1, 32, 1)

    def make():
        torch.manual_seed(0)
        m = DeepLab(5, "mobilenet", False, 16).set_compute_dtype(torch.float32)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        return m, torch.optim.SGD(m.parameters(), lr=1e-2)   # SGD: weight deltas stay proportional to gradient deltas

    # hand-rolled reference loop (the structure of utils_fit.py:46-121, fp32 branch)
    ref, opt = make()
    ref.train()
    ref_losses = []
    for imgs, pngs, labels in train:
        opt.zero_grad()
        y = ref(imgs)
        loss = Focal_Loss(y, pngs, torch.from_numpy(cls_w), num_classes=5) + Dice_loss(y, labels)
        loss.backward()
        opt.step()
        ref_losses.append(float(loss))

    model, opt = make()
    hist, ev = _History(), _Eval()
    fit_one_epoch(model, model, hist, ev, opt, 0, len(train), len(val), train, val, 1, False, True, True, cls_w, 5,
                  False, None, 5, str(tmp_path))
    assert ev.calls == [1] and len(hist.losses) == 1
    assert abs(hist.losses[0] - sum(ref_losses) / len(ref_losses)) < 1e-4 * abs(hist.losses[0])
    files = sorted(os.listdir(tmp_path))
    assert "best_epoch_weights.pth" in files and "last_epoch_weights.pth" in files
    assert any(f.startswith("ep001-loss") and f.endswith(".pth") and "-val_loss" in f for f in files)
    saved = torch.load(os.path.join(tmp_path, "last_epoch_weights.pth"))
    for k, v in model.state_dict().items():
        assert torch.equal(saved[k], v), k
    # same trajectory as the hand-rolled loop: tiny-batch batch-norm gradients are ill-conditioned,
    # so the two runs are compared by the direction and size of the total update, not element-wise
    init, _ = make()
    d_ref = torch.cat([(ref.state_dict()[k] - v).flatten() for k, v in init.state_dict().items() if v.dtype.is_floating_point and "running" not in k])
    d_got = torch.cat([(saved[k] - v).flatten() for k, v in init.state_dict().items() if v.dtype.is_floating_point and "running" not in k])
    cos = float(torch.dot(d_ref, d_got) / (d_ref.norm() * d_got.norm()))
    assert cos > 0.98 and abs(float(d_got.norm() / d_ref.norm()) - 1) < 0.05, (cos, float(d_got.norm() / d_ref.norm()))
    assert not model.training   # the validation phase leaves the model in eval mode, as the reference does


@pytest.mark.parametrize("size,mix", [((97, 61), 1), ((48, 80), 0), ((64, 64), 2)])
def test_predictor_matches_reference_postprocessing(size, mix):
    import cv2
    from PIL import Image
    from cervix_b200.deeplab import DeeplabV3
    from cervix_b200.nets.deeplabv3_plus import DeepLab
    from cervix_b200.utils.utils import cvtColor, preprocess_input, resize_image

    torch.manual_seed(3)
    net = DeepLab(5, "mobilenet", False, 16)
    for m in net.modules():   # make the logits vary over the image
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    state = {k: v.clone() for k, v in net.state_dict().items()}
    pred = DeeplabV3(model_path=state, backbone="mobilenet", input_shape=[64, 64], cuda=False, mix_type=mix,
                     compute_dtype=torch.float32)
    rng = np.random.RandomState(0)
    img = Image.fromarray(rng.randint(0, 255, (size[0], size[1], 3), dtype=np.uint8))

    # the reference's chain on the same network output
    image = cvtColor(img)
    oh, ow = np.array(image).shape[:2]
    canvas, nw, nh = resize_image(image, (64, 64))
    data = np.expand_dims(np.transpose(preprocess_input(np.array(canvas, np.float32)), (2, 0, 1)), 0)
    with torch.no_grad():
        pr = pred.net(torch.from_numpy(data))[0]
        pr = torch.softmax(pr.permute(1, 2, 0), dim=-1).numpy()
    pr = pr[(64 - nh) // 2:(64 - nh) // 2 + nh, (64 - nw) // 2:(64 - nw) // 2 + nw]
    pr = cv2.resize(pr, (ow, oh), interpolation=cv2.INTER_LINEAR).argmax(axis=-1)

    got = np.array(pred.get_miou_png(img))
    assert got.shape == (oh, ow) and got.dtype == np.uint8
    assert (got == pr).mean() >= 0.999
    out = pred.detect_image(img)
    assert out.size == (ow, oh)
    if mix == 1:
        assert (np.array(out) == np.array(pred.colors, np.uint8)[got]).all()
    assert pred.get_FPS(img, 1) > 0


def test_miou_metrics_match_numpy_reference():
    """fast_hist / per_class_iu / PA / precision / accuracy vs the reference formulas (utils_metrics.py:37-118)."""
    from cervix_b200.utils.utils_metrics import fast_hist, per_Accuracy, per_class_iu, per_class_PA_Recall, per_class_Precision
    rng = np.random.RandomState(0)
    n = 5
    gt = rng.randint(0, 6, (2, 37, 41)).astype(np.uint8)
    gt[gt == 5] = 255                       # ignore label, as in the VOC-style PNGs
    pred = rng.randint(0, 5, (2, 37, 41)).astype(np.uint8)
    a, b = gt.reshape(-1), pred.reshape(-1)
    k = (a >= 0) & (a < n)
    ref = np.bincount(n * a[k].astype(int) + b[k], minlength=n ** 2).reshape(n, n)
    got = fast_hist(a, b, n)
    assert got.dtype == np.int64 and (got == ref).all()
    acc = torch.zeros(n, n, dtype=torch.int64)
    for i in range(2):                       # accumulation over a set
        fast_hist(gt[i], pred[i], n, acc)
    assert (acc.numpy() == ref).all()
    assert np.allclose(per_class_iu(got), np.diag(ref) / np.maximum(ref.sum(1) + ref.sum(0) - np.diag(ref), 1))
    assert np.allclose(per_class_PA_Recall(got), np.diag(ref) / np.maximum(ref.sum(1), 1))
    assert np.allclose(per_class_Precision(got), np.diag(ref) / np.maximum(ref.sum(0), 1))
    assert np.isclose(per_Accuracy(got), np.diag(ref).sum() / max(ref.sum(), 1))
