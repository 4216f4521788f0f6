"""GPU parity tests of the tcgen05/TMEM/TMA implicit-GEMM convolution (forward, data gradient,
weight gradient) against the plain-torch specification, on the shapes of SURVEY.md appendix A
(incl. the ragged 728/304/48 channel counts, atrous taps that fall entirely into the padding,
and image sizes that do not divide the pixel tile)."""
import pytest
import torch

from cervix_b200.backend import ConvGeom, get_backend
from tests.emu_backend import EmuBackend
from tools.tc_probe import CASES

pytestmark = pytest.mark.gpu
EMU = EmuBackend()


@pytest.fixture(params=[1, 2, 4])
def pairs(request):
    """CTA pairs per cluster of the forward / data-gradient kernel, forced even for tiny problems so that the
    weight-multicast paths are exercised on every shape."""
    from cervix_b200 import _lib
    lib = _lib.load()
    assert lib.cvx_conv_tc_set_pairs(request.param, 1) == 0
    yield request.param
    assert lib.cvx_conv_tc_set_pairs(0, 0) == 0


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_tc_conv_case(idx, pairs):
    torch.backends.cudnn.allow_tf32 = False
    kind, n, h, w, cin, cout, k, pad, dil, bias = CASES[idx]
    if pairs != 1 and kind not in ("fwd", "dgrad"):
        pytest.skip("pairs per cluster only affects the stride-1 forward / data-gradient kernel")
    B = get_backend()
    g = ConvGeom(n, h, w, cin, cout, k, k, 2 if kind == "fwd_s2" else 1, pad, dil)
    gen = torch.Generator(device="cuda").manual_seed(idx)
    x = torch.randn((n, h, w, cin), generator=gen, device="cuda").bfloat16()
    wt = torch.randn((cout, cin, k, k), generator=gen, device="cuda") * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn((cout,), generator=gen, device="cuda") if bias else None
    dy = torch.randn((n, g.ho, g.wo, cout), generator=gen, device="cuda").bfloat16()
    if kind in ("fwd", "fwd_s2"):
        wp = B.pack_weight(wt, torch.bfloat16, False)
        got, ref, tol = B.conv_fwd(x, wp, b, g, True), EMU.conv_fwd(x, wp, b, g, False), 1.2e-2
    elif kind == "dgrad":
        wpt = B.pack_weight(wt, torch.bfloat16, True)
        got, ref, tol = B.conv_dgrad(dy, wpt, g, True), EMU.conv_dgrad(dy, wpt, g, False), 1.2e-2
    else:
        got, ref, tol = B.conv_wgrad(x, dy, g, True), EMU.conv_wgrad(x, dy, g, False), 3e-3
    err = float((got.float() - ref.float()).abs().max()) / float(ref.float().abs().max())
    assert err < tol, "%s case %d: rel err %.3e" % (kind, idx, err)
