"""Training augmentation on the device (reference: DeeplabDataset.get_random_data,
Segmentation/deeplabv3+/utils/dataloader.py:55-154).

CPU part: (1) every stage of the numpy oracle (oracle/augment_ref.py) against the installed Pillow / OpenCV, the libraries
the reference calls; (2) the oracle's composition against golden outputs of the reference's OWN function under fixed
numpy seeds (tests/golden/augment.npz, oracle/make_golden_augment.py); (3) the product's host side - random decisions,
integer tables, batch packing - through a numpy interpreter of the kernels (tests/emu_backend.py) against the same goldens.
GPU part: the CUDA kernels through the C ABI against the goldens, against the oracle at 512 x 512, and inside a train step.
Everything is 8-bit integer work: the bar is bit equality."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import augment_ref as A

cv2 = pytest.importorskip("cv2")
Image = pytest.importorskip("PIL.Image")


def _golden_cases(golden_dir):
    z = np.load(golden_dir + "/augment.npz")
    for i in range(int(z["n"])):
        yield dict(i=i, seed=int(z["seed_%d" % i]), shape=tuple(int(v) for v in z["shape_%d" % i]), random=bool(z["random_%d" % i]),
                   img=z["img_%d" % i], lab=z["lab_%d" % i], out_img=z["out_img_%d" % i], out_lab=z["out_lab_%d" % i])


def _smooth(rng, h, w):
    low = rng.randint(0, 256, (h // 5 + 2, w // 5 + 2, 3)).astype(np.uint8)
    return np.asarray(Image.fromarray(low).resize((w, h), Image.BILINEAR))


# ------------------------------------------------------------------------------------------------ oracle vs libraries
@pytest.mark.parametrize("ih,iw,oh,ow", [(37, 53, 64, 64), (100, 80, 33, 47), (64, 64, 128, 31), (50, 70, 50, 140),
                                         (300, 200, 77, 131), (31, 17, 31, 64)])
def test_oracle_bicubic_and_nearest_equal_pillow(ih, iw, oh, ow):
    rng = np.random.RandomState(ih * 1000 + ow)
    for src in (rng.randint(0, 256, (ih, iw, 3)).astype(np.uint8), _smooth(rng, ih, iw)):
        ref = np.asarray(Image.fromarray(src).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(A.pil_resize_u8(src, ow, oh), ref)
    lab = rng.randint(0, 6, (ih, iw)).astype(np.uint8)
    assert np.array_equal(A.pil_resize_nearest(lab, ow, oh), np.asarray(Image.fromarray(lab).resize((ow, oh), Image.NEAREST)))


def test_oracle_blur_and_rotation_equal_opencv():
    rng = np.random.RandomState(3)
    for h, w in ((64, 64), (33, 47), (5, 9), (96, 64)):
        src = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        assert np.array_equal(A.cv_gaussian5_u8(src), cv2.GaussianBlur(src, (5, 5), 0))
    for h, w in ((64, 64), (48, 80)):
        src = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        lab = rng.randint(0, 6, (h, w)).astype(np.uint8)
        for rot in range(-10, 11):
            m = cv2.getRotationMatrix2D((w // 2, h // 2), -rot, scale=1)
            assert np.array_equal(A.cv_warp_cubic_u8(src, rot, 128),
                                  cv2.warpAffine(src, m, (w, h), flags=cv2.INTER_CUBIC, borderValue=(128, 128, 128))), rot
            assert np.array_equal(A.cv_warp_nearest_u8(lab, rot, 0),
                                  cv2.warpAffine(lab, m, (w, h), flags=cv2.INTER_NEAREST, borderValue=(0))), rot


def test_oracle_hsv_equals_opencv_on_every_input():
    """RGB -> HSV for all 2^24 colours; HSV -> RGB for all 180 x 256 x 256 triples in OpenCV's vector loop and, on ragged
    widths, in its scalar tail (the two differ in rounding - see oracle/augment_ref.py)."""
    g = np.arange(256, dtype=np.uint8)
    for r0 in range(0, 256, 64):           # four slabs of 64 x 256 x 256 colours
        rgb = np.stack(np.meshgrid(g[r0:r0 + 64], g, g, indexing="ij"), -1).reshape(64 * 16, 4096, 3)
        assert np.array_equal(A.cv_rgb2hsv_u8(rgb), cv2.cvtColor(rgb, cv2.COLOR_RGB2HSV))
    for h0 in range(0, 180, 60):
        hsv = np.stack(np.meshgrid(np.arange(h0, h0 + 60, dtype=np.uint8), g, g, indexing="ij"), -1).reshape(60 * 16, 4096, 3)
        assert np.array_equal(A.cv_hsv2rgb_u8(hsv), cv2.cvtColor(hsv, cv2.COLOR_HSV2RGB))
    rng = np.random.RandomState(1)
    for w in (70, 100, 37, 7):
        hsv = np.stack([rng.randint(0, 180, (200, w)), rng.randint(0, 256, (200, w)), rng.randint(0, 256, (200, w))], -1).astype(np.uint8)
        assert np.array_equal(A.cv_hsv2rgb_u8(hsv), cv2.cvtColor(hsv, cv2.COLOR_HSV2RGB)), w


def test_oracle_reproduces_the_reference_function(golden_dir):
    """Goldens = outputs of the reference's get_random_data under np.random.seed(seed): same decisions, same pixels."""
    n = 0
    for c in _golden_cases(golden_dir):
        ih, iw = c["lab"].shape
        if c["random"]:
            np.random.seed(c["seed"])
            p = A.draw_params(iw, ih, c["shape"], np.random)
        else:
            p = A.letterbox_params(iw, ih, c["shape"])
        img, lab = A.apply_params(c["img"], c["lab"], c["shape"], p)
        assert np.array_equal(img, c["out_img"]) and np.array_equal(lab, c["out_lab"]), c["i"]
        n += 1
    assert n == 10


# ------------------------------------------------------------------------------------------------ host side of the product
def test_host_tables_equal_the_oracle():
    from cervix_b200.multimodal.pil_resample import resample_tables
    from cervix_b200.utils import dataloader as D
    rng = np.random.RandomState(0)
    for _ in range(25):
        a, b = int(rng.randint(5, 900)), int(rng.randint(5, 700))
        lo, cnt, k = resample_tables(a, b, "bicubic")
        lo2, cnt2, k2 = A.pil_coeffs(a, b, "bicubic")
        assert np.array_equal(lo, lo2) and np.array_equal(cnt, cnt2) and np.array_equal(k, k2), (a, b)
        assert np.array_equal(D.nearest_table(a, b), A.pil_nearest_index(a, b)), (a, b)
    assert np.array_equal(D.cubic_weights().reshape(32, 32, 4, 4), A.cubic_table())
    for w, h in ((96, 96), (128, 64), (512, 512)):
        for rot in (-10, -3, 0, 1, 7, 10):
            t = D.rotation_tables(w, h, rot).astype(np.int64)
            ad, bd, x0, y0 = t[:w], t[w:2 * w], t[2 * w:2 * w + h], t[2 * w + h:]
            X, Y, fx, fy = A.warp_coords(A.rotation_matrix(w, h, rot), w, h, nearest=False)
            Xp, Yp = (x0[:, None] + 16 + ad[None, :]) >> 5, (y0[:, None] + 16 + bd[None, :]) >> 5
            assert np.array_equal(Xp >> 5, X) and np.array_equal(Yp >> 5, Y) and np.array_equal(Xp & 31, fx) and np.array_equal(Yp & 31, fy)
            Xn, Yn, _, _ = A.warp_coords(A.rotation_matrix(w, h, rot), w, h, nearest=True)
            assert np.array_equal((x0[:, None] + 512 + ad[None, :]) >> 10, Xn) and np.array_equal((y0[:, None] + 512 + bd[None, :]) >> 10, Yn)
    r = np.array([1.07, 0.4, 1.29])
    assert np.array_equal(D.hsv_luts(r), np.concatenate(A.hsv_luts(r)))


def test_random_decisions_follow_the_reference_order():
    from cervix_b200.utils import dataloader as D
    for seed in range(40):
        np.random.seed(seed)
        p = D.draw_params(83, 57, (96, 96))
        tail = np.random.rand()                      # the generator must be left in the same state as well
        np.random.seed(seed)
        q = A.draw_params(83, 57, (96, 96), np.random)
        assert tail == np.random.rand()
        assert all(np.array_equal(p[k], q[k]) for k in q), seed
    assert D.letterbox_params(83, 57, (96, 64)) == {**A.letterbox_params(83, 57, (96, 64))}


def test_descriptor_layout_matches_the_header():
    from cervix_b200 import _lib
    from cervix_b200.utils.dataloader import AUG_SAMPLE
    assert AUG_SAMPLE.itemsize == C.sizeof(_lib.AugSample) == 96
    for name, _ in _lib.AugSample._fields_:
        assert AUG_SAMPLE.fields[name][1] == getattr(_lib.AugSample, name).offset, name
    with open(__file__.rsplit("/tests/", 1)[0] + "/include/cervix_b200.h") as f:
        header = f.read()
    body = header[header.index("typedef struct cvx_aug_sample {"):header.index("} cvx_aug_sample;")]
    for name, _ in _lib.AugSample._fields_:
        assert name in body, name


def _run_cases(golden_dir, device):
    from cervix_b200.utils.dataloader import DeeplabDataset
    ds = DeeplabDataset(["x"], (96, 96), 5, True, "/nonexistent")
    n = 0
    for c in _golden_cases(golden_dir):
        np.random.seed(c["seed"] if c["random"] else 0)
        img, lab = ds.get_random_data(Image.fromarray(c["img"]), Image.fromarray(c["lab"]), c["shape"], random=c["random"],
                                      device=device)
        assert img.dtype == np.uint8 and img.shape == c["shape"] + (3,) and lab.shape == c["shape"]
        assert np.array_equal(lab, c["out_lab"]), ("label", c["i"])
        assert np.array_equal(img, c["out_img"]), ("image", c["i"], int((img != c["out_img"]).sum()))
        n += 1
    return n


def test_host_packing_through_the_kernel_interpreter(golden_dir):
    """DeeplabDataset.get_random_data of the drop-in with the emulated ABI: decisions, tables and packing are the
    product's, the arithmetic is the numpy interpreter of the kernels."""
    from cervix_b200 import backend
    from tests.emu_backend import EmuBackend
    prev = backend.set_backend(EmuBackend())
    try:
        assert _run_cases(golden_dir, "cpu") == 10
    finally:
        backend.set_backend(prev)


def test_collate_packs_mixed_sizes_and_rejects_bad_input():
    from cervix_b200.utils import dataloader as D
    rng = np.random.RandomState(2)
    ds = D.DeeplabDataset(["a", "b"], (64, 96), 5, True, "/nonexistent")
    np.random.seed(4)
    items = [ds.decode(Image.fromarray(rng.randint(0, 256, (h, w, 3)).astype(np.uint8)), Image.fromarray(rng.randint(0, 6, (h, w)).astype(np.uint8)))
             for h, w in ((40, 50), (70, 30), (40, 50))]
    plan = D.deeplab_dataset_collate(items)
    assert len(plan) == 3 and plan.shape == (64, 96) and plan.samples.shape == (3, 96) and plan.luts.numel() == 3 * 768
    rec = np.frombuffer(plan.samples.numpy().tobytes(), dtype=D.AUG_SAMPLE)
    assert all(int(o) % 16 == 0 for o in rec["src_off"]) and rec["lut"].tolist() == [0, 768, 1536]
    assert plan.max_elems == max(int(r["ih"]) * int(r["nw"]) for r in rec if r["iw"] != r["nw"])
    val = D.DeeplabDataset(["a"], (64, 96), 5, False, "/nonexistent").decode(Image.fromarray(items[0][0]), Image.fromarray(items[0][1]))
    assert val[2]["r"] is None and D.pack_batch([val[:3]], (64, 96)).samples.shape == (1, 96)
    with pytest.raises(TypeError):
        D.pack_batch([(items[0][0].astype(np.float32), items[0][1], items[0][2])], (64, 96))
    with pytest.raises(ValueError):
        D.pack_batch([(items[0][0], items[1][1], items[0][2])], (64, 96))
    with pytest.raises(ValueError):
        D.pack_batch([(items[0][0], items[0][1], dict(items[0][2], nw=0))], (64, 96))


# ------------------------------------------------------------------------------------------------ CUDA
@pytest.mark.gpu
def test_gpu_augmentation_equals_the_reference_function(golden_dir):
    assert _run_cases(golden_dir, "cuda") == 10


@pytest.mark.gpu
def test_gpu_batched_augmentation_equals_the_oracle_at_full_size():
    """One launch set for a batch of differently sized sources on the 512 x 512 canvas (BASELINE configs[2] input size),
    every stage switched on somewhere, against the numpy oracle."""
    from cervix_b200.utils import dataloader as D
    rng = np.random.RandomState(7)
    items, want = [], []
    forced = [dict(blur=True, rotate=True), dict(blur=False, rotate=True), dict(blur=True, rotate=False), dict()]
    for k, (ih, iw) in enumerate(((375, 500), (600, 420), (256, 256), (512, 300))):
        img, lab = _smooth(rng, ih, iw), rng.randint(0, 7, (ih, iw)).astype(np.uint8)
        np.random.seed(100 + k)
        p = D.draw_params(iw, ih, (512, 512))
        p.update(forced[k])
        if p["rotate"] and p["rotation"] == 0:
            p["rotation"] = 7 - 3 * k
        items.append((img, lab, p))
        want.append(A.apply_params(img, lab, (512, 512), p))
    val = D.letterbox_params(500, 375, (512, 512))
    items.append((items[0][0], items[0][1], val))
    want.append(A.apply_params(items[0][0], items[0][1], (512, 512), val))
    aug = D.DeviceAugmenter("cuda")
    imgs, labs = aug.run(D.pack_batch(items, (512, 512)))
    assert imgs.shape == (5, 512, 512, 3) and imgs.dtype == torch.uint8 and labs.shape == (5, 512, 512)
    for k, (wi, wl) in enumerate(want):
        assert np.array_equal(labs[k].cpu().numpy(), wl), k
        assert np.array_equal(imgs[k].cpu().numpy(), wi), (k, int((imgs[k].cpu().numpy() != wi).sum()))


@pytest.mark.gpu
def test_gpu_hsv_jitter_on_a_ragged_canvas_and_all_colours():
    """Identity geometry (source = canvas size), so the output is the HSV stage alone: every RGB colour once with unit
    gains (the conversion round trip), and a canvas width with a scalar tail."""
    from cervix_b200.utils import dataloader as D
    aug = D.DeviceAugmenter("cuda")
    g = np.arange(256, dtype=np.uint8)
    rgb = np.stack(np.meshgrid(g[:64], g, g, indexing="ij"), -1).reshape(1024, 4096, 3)
    for r in (np.array([1.0, 1.0, 1.0]), np.array([1.08, 0.45, 1.21])):
        p = dict(nw=4096, nh=1024, flip=False, dx=0, dy=0, blur=False, rotate=False, rotation=0, r=r)
        out, _ = aug.run(D.pack_batch([(rgb, np.zeros((1024, 4096), np.uint8), p)], (1024, 4096)))
        assert np.array_equal(out[0].cpu().numpy(), A.hsv_jitter_u8(rgb, r))
    rng = np.random.RandomState(5)
    src = rng.randint(0, 256, (50, 70, 3)).astype(np.uint8)
    p = dict(nw=70, nh=50, flip=True, dx=0, dy=0, blur=False, rotate=False, rotation=0, r=np.array([0.93, 1.6, 0.8]))
    out, _ = aug.run(D.pack_batch([(src, np.zeros((50, 70), np.uint8), p)], (50, 70)))
    assert np.array_equal(out[0].cpu().numpy(), A.hsv_jitter_u8(src[:, ::-1], p["r"]))


@pytest.mark.gpu
def test_gpu_device_loader_feeds_the_train_step():
    """DeviceAugmentLoader over collated batches: same pixels as a direct run, and the uint8 batches go straight into
    SegTrainer.step (the /255, the ignore-label clamp and the one-hot target happen inside the step's kernels)."""
    from cervix_b200.engine import SegTrainer
    from cervix_b200.nets.deeplabv3_plus import DeepLab
    from cervix_b200.utils import dataloader as D
    rng = np.random.RandomState(11)
    ds = D.DeeplabDataset(["x"] * 8, (64, 64), 5, True, "/nonexistent")
    np.random.seed(21)
    decoded = [ds.decode(Image.fromarray(_smooth(rng, 50 + 7 * i, 90 - 5 * i)), Image.fromarray(rng.randint(0, 7, (50 + 7 * i, 90 - 5 * i)).astype(np.uint8)))
               for i in range(8)]
    plans = [D.deeplab_dataset_collate(decoded[i:i + 4]).pin_memory() for i in (0, 4)]
    direct = [D.DeviceAugmenter("cuda").run(p) for p in plans]
    torch.manual_seed(0)
    model = DeepLab(5, "mobilenet", False, 16).set_compute_dtype(torch.float32).cuda().train()
    tr = SegTrainer(model, lr=1e-3, cls_weights=[1, 1, 5, 3, 4])
    n = 0
    for (imgs, labs, onehot), (di, dl) in zip(D.DeviceAugmentLoader(plans, "cuda"), direct):
        assert onehot is None and torch.equal(imgs, di) and torch.equal(labs, dl)
        out = tr.step(imgs, labs)
        assert torch.isfinite(out).all()
        n += 1
    assert n == 2
    want = [A.apply_params(img, lab, (64, 64), p) for img, lab, p, _ in decoded[:4]]
    assert all(np.array_equal(direct[0][0][k].cpu().numpy(), want[k][0]) and np.array_equal(direct[0][1][k].cpu().numpy(), want[k][1]) for k in range(4))


@pytest.mark.gpu
def test_gpu_fit_one_epoch_takes_the_drop_in_dataloader(tmp_path):
    """train.py's loader construction with the drop-in names - DataLoader(DeeplabDataset(...), collate_fn=
    deeplab_dataset_collate) over VOC-layout files (train.py:503-508) - handed to fit_one_epoch unchanged: the batches
    are packed plans, fit_one_epoch routes them through the device augmentation; uint8 batches, implicit one-hot labels,
    eager steps then the captured graph; the validation loader takes the letterbox path."""
    from torch.utils.data import DataLoader
    from cervix_b200.utils import dataloader as D
    from cervix_b200.utils.utils_fit import fit_one_epoch
    from tests.test_engine_gpu import _Ev, _Hist, _small_model
    rng = np.random.RandomState(13)
    (tmp_path / "VOC2007" / "JPEGImages").mkdir(parents=True)
    (tmp_path / "VOC2007" / "SegmentationClass").mkdir(parents=True)
    names = []
    for i in range(24):
        h, w = 48 + 3 * (i % 7), 80 - 4 * (i % 5)
        Image.fromarray(_smooth(rng, h, w)).save(tmp_path / "VOC2007" / "JPEGImages" / ("im%02d.jpg" % i), quality=92)
        Image.fromarray(rng.randint(0, 7, (h, w)).astype(np.uint8)).save(tmp_path / "VOC2007" / "SegmentationClass" / ("im%02d.png" % i))
        names.append("im%02d\n" % i)
    np.random.seed(5)
    train = DataLoader(D.DeeplabDataset(names, (64, 64), 5, True, str(tmp_path)), shuffle=False, batch_size=4, num_workers=0,
                       pin_memory=True, drop_last=True, collate_fn=D.deeplab_dataset_collate)
    val = DataLoader(D.DeeplabDataset(names[:4], (64, 64), 5, False, str(tmp_path)), shuffle=False, batch_size=4, num_workers=0,
                     pin_memory=True, drop_last=True, collate_fn=D.deeplab_dataset_collate)
    model = _small_model()
    opt = torch.optim.Adam(model.parameters(), 3e-4)
    hist, ev = _Hist(), _Ev()
    save = tmp_path / "logs"
    save.mkdir()
    fit_one_epoch(model, model, hist, ev, opt, 0, len(train), len(val), train, val, 1, True, True, True,
                  np.array([1, 1, 5, 3, 4], np.float32), 5, False, None, 1, str(save))
    tr = model._cvx_trainer
    assert tr.t == 6 and tr.graph is not None and len(opt.state) == 0
    assert len(hist.losses) == 1 and np.isfinite(hist.losses[0]) and np.isfinite(hist.val_loss[0])
    # what the validation loader delivered is the reference's letterbox of the decoded file
    plan = next(iter(val))
    imgs, labs = D.DeviceAugmenter("cuda").run(plan)
    src = np.asarray(Image.open(tmp_path / "VOC2007" / "JPEGImages" / "im00.jpg"))
    lab = np.asarray(Image.open(tmp_path / "VOC2007" / "SegmentationClass" / "im00.png"))
    wi, wl = A.apply_params(src, lab, (64, 64), A.letterbox_params(src.shape[1], src.shape[0], (64, 64)))
    assert np.array_equal(imgs[0].cpu().numpy(), wi) and np.array_equal(labs[0].cpu().numpy(), wl)


def test_fit_one_epoch_loader_hooks_without_a_device():
    """Host logic of the drop-in loader inside utils_fit: a DataLoader built with the drop-in collate function is
    recognised (and refused loudly when there is no CUDA device - the augmentation has no CPU fallback), any other loader
    passes through; uint8 batches with implicit one-hot labels are finished on whatever device they live on."""
    from torch.utils.data import DataLoader
    from cervix_b200.utils import dataloader as D
    from cervix_b200.utils import utils_fit as F
    ds = D.DeeplabDataset(["a"], (32, 32), 5, True, "/nonexistent")
    ours = DataLoader(ds, batch_size=1, collate_fn=D.deeplab_dataset_collate)
    other = [(torch.zeros(1, 3, 8, 8), torch.zeros(1, 8, 8, dtype=torch.long), torch.zeros(1, 8, 8, 6))]
    assert F._wrap_loader(other, False, 0) is other
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            F._wrap_loader(ours, True, 0)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            F._wrap_loader(ours, False, 0)
    pngs = torch.tensor([[[0, 4, 5, 255]]], dtype=torch.uint8)
    imgs, p, lab, w = F._to_device((torch.zeros(1, 1, 4, 3, dtype=torch.uint8), pngs, None), np.ones(5, np.float32), False, 0, 5)
    assert p.dtype == torch.int64 and p.tolist() == [[[0, 4, 5, 5]]]
    assert lab.shape == (1, 1, 4, 6) and lab.argmax(-1).tolist() == [[[0, 4, 5, 5]]]
