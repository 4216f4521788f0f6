"""ResNet-101 patch encoder (classifier image branches) against torchvision's
resnet101 - the architecture the reference instantiates - on identical random weights and inputs.
CPU: host graph on the emulation backend.  GPU: the CUDA engine in fp32 (1e-3) and bf16 (cosine)."""
import pytest
import torch
import torchvision

import cervix_b200.backend as backend
from cervix_b200.multimodal.patch_encoder import ResNet101Encoder, extract_features, split_patches
from tests.emu_backend import EmuBackend


def _pair(seed=0):
    torch.manual_seed(seed)
    ref = torchvision.models.resnet101(weights=None)
    ref.fc = torch.nn.Linear(2048, 1024)
    for m in ref.modules():   # non-trivial running statistics, damped residual branches (keeps activations O(1))
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.3, 0.6); m.bias.data.normal_(0, 0.05)
    ref.eval()
    enc = ResNet101Encoder()
    assert list(enc.state_dict().keys()) == list(ref.state_dict().keys())
    enc.load_state_dict(ref.state_dict(), strict=True)
    return ref, enc.eval()


def test_cpu_host_graph_matches_torchvision():
    prev = backend.set_backend(EmuBackend())
    try:
        ref, enc = _pair()
        enc.set_compute_dtype(torch.float32)
        x = torch.rand(2, 3, 64, 96)
        with torch.no_grad():
            a, b = enc(x), ref(x)
        assert float((a - b).abs().max()) < 1e-4 * float(b.abs().max())
    finally:
        backend.set_backend(prev)


def test_folded_bottleneck_matches_unfolded_on_emulation():
    """The inference fast path (BatchNorm folded into the conv weights, bias + residual + ReLU in the conv epilogue)
    against the operator-by-operator eval forward of the same block, on the emulated C ABI."""
    from cervix_b200.multimodal.patch_encoder import Bottleneck
    prev = backend.set_backend(EmuBackend())
    try:
        torch.manual_seed(0)
        for inpl, planes, stride in ((64, 16, 1), (64, 32, 2)):
            down = None
            if stride != 1 or inpl != planes * 4:
                down = torch.nn.Sequential(torch.nn.Conv2d(inpl, planes * 4, 1, stride, bias=False),
                                           torch.nn.BatchNorm2d(planes * 4))
            blk = Bottleneck(inpl, planes, stride, down).eval()
            for m in blk.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5)
                    m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
            x = torch.randn(2, 12, 10, inpl)            # NHWC fp32: the emulation keeps full precision
            with torch.no_grad():
                want = blk(x)
                got = blk.forward_folded(x, torch.ones(2048))
            assert got.shape == want.shape
            assert float((got - want).abs().max()) < 2e-2 * float(want.abs().max())   # weights are packed in bf16
    finally:
        backend.set_backend(prev)


def test_split_patches_order_and_normalisation():
    prev = backend.set_backend(EmuBackend())
    try:
        img = torch.rand(1, 3, 1024, 1024)
        p = split_patches(img)
    finally:
        backend.set_backend(prev)
    assert p.shape == (16, 3, 256, 256)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1); std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    # reference loop order: for i in x-blocks: for j in y-blocks -> patch index = i*4 + j
    for i, j in ((0, 0), (1, 0), (0, 3), (2, 1)):
        want = (img[0, :, j * 256:(j + 1) * 256, i * 256:(i + 1) * 256] - mean) / std
        assert torch.allclose(p[i * 4 + j], want, atol=1e-5)


def _split_spec(img, new_size, patch):
    """What cvx_split_patches must produce, from stock torch ops on the CPU (NHWC fp32)."""
    return EmuBackend().split_patches(img.cpu(), new_size, patch, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225),
                                      torch.float32)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,new_size,patch", [((2, 3, 512, 512), 1024, 256), ((1, 3, 300, 280), 1024, 256),
                                                  ((1, 3, 1024, 1024), 1024, 256), ((1, 3, 1500, 1210), 1024, 256),
                                                  ((3, 3, 96, 80), 256, 64), ((1, 3, 37, 53), 64, 16)])
def test_gpu_split_patches_kernel(shape, new_size, patch):
    """The one-launch resize + x-major split + normalise kernel against F.interpolate + slicing, fp32 and bf16."""
    from cervix_b200.multimodal.patch_encoder import split_patches_nhwc
    torch.manual_seed(sum(shape))
    img = torch.rand(*shape)
    want = _split_spec(img, new_size, patch)
    got = split_patches_nhwc(img.cuda(), new_size, patch, torch.float32).cpu()
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 2e-5
    got16 = split_patches_nhwc(img.cuda(), new_size, patch, torch.bfloat16)
    assert got16.dtype == torch.bfloat16
    assert float((got16.float().cpu() - want).abs().max()) < 2e-2          # bf16 storage of values in [-2.2, 2.7]
    nchw = split_patches(img.cuda(), new_size, patch).cpu()
    assert torch.equal(nchw, got.permute(0, 3, 1, 2))


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,new_size,patch", [(512, 512, 1024, 256), (300, 417, 1024, 256), (1500, 2000, 1024, 256),
                                                (1024, 1024, 1024, 256), (2049, 1025, 1024, 256), (97, 33, 64, 16)])
def test_gpu_split_patches_u8_equals_pillow(h, w, new_size, patch):
    """the reference resizes the DECODED uint8 image with ``img.resize((1024, 1024), Image.BILINEAR)``
    (Graph_Structure(data_augmentation).py:154) - Pillow's fixed-point, two-pass, antialiasing-when-reducing resampler, not
    ``F.interpolate``.  The uint8 kernel must reproduce the installed Pillow BIT FOR BIT, enlarging (512 -> 1024) and
    reducing (2000 -> 1024) alike, through the x-major crops, ToTensor and Normalize."""
    import numpy as np
    from PIL import Image
    from cervix_b200.multimodal.patch_encoder import IMAGENET_MEAN, IMAGENET_STD, split_patches_nhwc
    rng = np.random.RandomState(h * 7 + w)
    imgs = rng.randint(0, 256, (2, h, w, 3), dtype=np.uint8)
    mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
    want = []
    for a in imgs:
        resized = Image.fromarray(a).resize((new_size, new_size), Image.BILINEAR)
        for i in range(0, new_size, patch):                  # the reference's loop order: outer x, inner y
            for j in range(0, new_size, patch):
                t = torch.from_numpy(np.asarray(resized.crop((i, j, i + patch, j + patch)))).permute(2, 0, 1).float().div(255)
                want.append(t.sub(mean).div(std))
    want = torch.stack(want)
    got = split_patches_nhwc(torch.from_numpy(imgs).cuda(), new_size, patch, torch.float32).permute(0, 3, 1, 2).cpu()
    assert got.shape == want.shape
    assert torch.equal(got, want), float((got - want).abs().max())
    got16 = split_patches_nhwc(torch.from_numpy(imgs).cuda(), new_size, patch, torch.bfloat16)
    assert torch.equal(got16.float().cpu().permute(0, 3, 1, 2), want.bfloat16().float())


def test_pillow_tables_reproduce_pillow_on_the_host():
    """The coefficient tables the kernel consumes (multimodal/pil_resample.py), applied in numpy, against Pillow."""
    import numpy as np
    from PIL import Image
    from cervix_b200.multimodal.pil_resample import resize_u8_reference
    rng = np.random.RandomState(0)
    for h, w, o in ((128, 128, 256), (75, 104, 256), (375, 500, 256), (256, 160, 64)):
        a = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(a).resize((o, o), Image.BILINEAR))
        assert np.array_equal(resize_u8_reference(a, o, o), ref), (h, w, o)


@pytest.mark.gpu
def test_gpu_encode_images_equals_split_then_forward():
    _, enc = _pair(4)
    enc.set_compute_dtype(torch.bfloat16).cuda()
    img = torch.rand(2, 3, 128, 160).cuda()
    with torch.no_grad():
        a = enc.encode_images(img, 256, 64)
        b = enc(split_patches(img, 256, 64))
    assert a.shape == (32, 1024)
    assert float((a - b).abs().max()) < 2e-2 * float(b.abs().max())


@pytest.mark.gpu
def test_gpu_finish_batch_u8():
    """uint8 loader tail (dataloader.py:40-42 + preprocess_input) against numpy."""
    import numpy as np
    B = backend.get_backend()
    rng = np.random.RandomState(0)
    for n, h, w in ((2, 64, 48), (1, 7, 5), (3, 33, 31)):
        img = rng.randint(0, 256, (n, h, w, 3), dtype=np.uint8)
        lab = rng.randint(0, 9, (n, h, w), dtype=np.uint8)
        lab[0, 0, 0] = 255
        x, t = B.finish_batch_u8(torch.from_numpy(img).cuda(), torch.from_numpy(lab).cuda(), 5, torch.float32)
        want_t = lab.astype(np.int64); want_t[want_t >= 5] = 5
        assert np.array_equal(t.cpu().numpy(), want_t)
        assert float((x.cpu() - torch.from_numpy(img.astype(np.float32) / 255.0)).abs().max()) < 1e-7
        xb, _ = B.finish_batch_u8(torch.from_numpy(img).cuda(), None, 5, torch.bfloat16)
        assert torch.equal(xb.cpu(), torch.from_numpy(img.astype(np.float32) / 255.0).bfloat16())


@pytest.mark.gpu
def test_gpu_fp32_matches_torchvision():
    ref, enc = _pair(1)
    enc.set_compute_dtype(torch.float32).cuda()
    x = split_patches(torch.rand(1, 3, 300, 280).cuda())[:6].cpu()
    with torch.no_grad():
        b = ref(x)
        a = extract_features(x.cuda(), enc).cpu()
    assert a.shape == (6, 1024)
    assert float((a - b).abs().max()) < 1e-3 * float(b.abs().max())


@pytest.mark.gpu
def test_gpu_bf16_tracks_torchvision():
    ref, enc = _pair(2)
    enc.set_compute_dtype(torch.bfloat16).cuda()
    x = split_patches(torch.rand(1, 3, 256, 256).cuda()).cpu()
    with torch.no_grad():
        b = ref(x)
        a = extract_features(x.cuda(), enc).cpu()
    cos = torch.nn.functional.cosine_similarity(a, b, dim=1)
    assert float(cos.min()) > 0.995, float(cos.min())
    assert float((a - b).abs().max()) < 0.08 * float(b.abs().max())


def test_age_node_features_match_reference_recipe():
    """x_cli rows against a numpy restatement of Graph_Structure(...).py:70-127."""
    import numpy as np
    from cervix_b200.multimodal.cli_features import AgeNodeFeatures, age_to_one_hot
    ages = [23, 35, 41, 58, 64, 79, 30, 52]
    torch.manual_seed(0)
    mod = AgeNodeFeatures(max_age=max(ages))
    got = mod(ages)
    assert got.shape == (len(ages), 4, 1024)
    hi, lo = max(ages), min(ages)
    for g, age in enumerate(ages):
        norm = (age - (hi + lo) / 2) / (hi - lo) * 2
        assert np.array_equal(got[g, 0].numpy(), age_to_one_hot(age).astype(np.float32))
        assert np.array_equal(got[g, 1].numpy(), age_to_one_hot(norm).astype(np.float32))   # bin 0 or the LAST bin
        assert torch.equal(got[g, 2], mod.age_table[age])
        assert torch.equal(got[g, 3], mod.age_std_table[int((norm + 1) / 2 * 100)])
    assert got[ages.index(23), 1].argmax() == 19 and got[ages.index(79), 1].argmax() == 0


def test_pil_entry_points_match_the_batched_path():
    """resize_and_split_image / extract_features in the reference's PIL-in numpy-out form (Graph_Structure...py:151-168)
    against the tensor path, on the emulated C ABI."""
    import numpy as np
    from PIL import Image
    from cervix_b200.multimodal.patch_encoder import extract_patient_features, resize_and_split_image
    prev = backend.set_backend(EmuBackend())
    try:
        _, enc = _pair(3)
        enc.set_compute_dtype(torch.float32)
        rng = np.random.RandomState(0)
        img = Image.fromarray(rng.randint(0, 255, (96, 80, 3), dtype=np.uint8))
        patches = resize_and_split_image(img, 256, 64)       # small sizes keep the CPU test quick
        assert len(patches) == 16 and patches[0].size == (64, 64)
        resized = np.asarray(img.resize((256, 256), Image.BILINEAR))
        assert np.array_equal(np.asarray(patches[1]), resized[64:128, 0:64])       # index 1 = (x block 0, y block 1)
        assert np.array_equal(np.asarray(patches[4]), resized[0:64, 64:128])       # index 4 = (x block 1, y block 0)
        one = extract_features(patches[5], enc)
        assert isinstance(one, np.ndarray) and one.shape == (1024,)
        import torchvision.transforms as T
        tf = T.Compose([T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        same = extract_features(patches[5], enc, tf, torch.device("cpu"))
        assert np.allclose(one, same, rtol=1e-4, atol=1e-5)
        batch = torch.stack([tf(p) for p in patches[4:7]])
        assert np.allclose(extract_features(batch, enc)[1].numpy(), one, rtol=1e-3, atol=1e-4)
    finally:
        backend.set_backend(prev)
