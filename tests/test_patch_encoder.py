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
    img = torch.rand(1, 3, 1024, 1024)
    p = split_patches(img)
    assert p.shape == (16, 3, 256, 256)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1); std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    # reference loop order: for i in x-blocks: for j in y-blocks -> patch index = i*4 + j
    for i, j in ((0, 0), (1, 0), (0, 3), (2, 1)):
        want = (img[0, :, j * 256:(j + 1) * 256, i * 256:(i + 1) * 256] - mean) / std
        assert torch.allclose(p[i * 4 + j], want, atol=1e-5)


@pytest.mark.gpu
def test_gpu_fp32_matches_torchvision():
    ref, enc = _pair(1)
    enc.set_compute_dtype(torch.float32).cuda()
    x = split_patches(torch.rand(1, 3, 300, 280))[:6]
    with torch.no_grad():
        b = ref(x)
        a = extract_features(x.cuda(), enc).cpu()
    assert a.shape == (6, 1024)
    assert float((a - b).abs().max()) < 1e-3 * float(b.abs().max())


@pytest.mark.gpu
def test_gpu_bf16_tracks_torchvision():
    ref, enc = _pair(2)
    enc.set_compute_dtype(torch.bfloat16).cuda()
    x = split_patches(torch.rand(1, 3, 256, 256))
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
