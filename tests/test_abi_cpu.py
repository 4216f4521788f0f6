"""CPU tests: the C-ABI library builds for sm_100a, loads, and exports exactly the symbols that
include/cervix_b200.h declares (no compute calls - there is no GPU here)."""
import os
import re
import subprocess

import pytest

import __graft_entry__ as entry
from cervix_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    entry.build()
    return _lib.load()


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "cervix_b200.h")).read()
    return sorted(set(re.findall(r"CVX_API\s+[\w\s\*]+?\b(cvx_\w+)\s*\(", text)))


def test_header_and_binding_agree(built):
    syms = _header_symbols()
    assert len(syms) >= 40
    assert sorted(_lib.PROTOTYPES) == syms
    for name in syms:
        assert hasattr(built, name), name


def test_library_exports_exactly_the_header(built):
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("cvx_"))
    assert exported == _header_symbols()


def test_abi_version_and_error_string(built):
    assert built.cvx_abi_version() == 1
    assert isinstance(built.cvx_last_error(), bytes)


def test_argument_validation_without_gpu(built):
    # NULL pointers are rejected before any CUDA call is made
    assert built.cvx_relu_fwd(None, None, 16, _lib.F32, None) == -1
    assert b"relu_fwd" in built.cvx_last_error()
    d = _lib.ConvDesc(1, 8, 8, 8, 8, 3, 3, 1, 1, 1, 7, 8, _lib.BF16)  # wrong ho
    import ctypes as C
    assert built.cvx_conv_fwd(C.byref(d), 1, 1, None, 1, None) == -1
    assert b"inconsistent" in built.cvx_last_error()
    d = _lib.ConvDesc(1, 8, 8, 3, 8, 3, 3, 1, 1, 1, 8, 8, _lib.BF16)  # C_in = 3 -> not a tensor-core shape
    assert built.cvx_conv_fwd_tc(C.byref(d), 1, 1, None, 1, None) == _lib.EUNSUPPORTED


def test_sass_contains_blackwell_tensor_path(built):
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.check_output([cuobjdump, "-sass", _lib.LIB_PATH]).decode()
    assert "UTCHMMA" in sass      # tcgen05.mma
    assert "UTMALDG" in sass      # TMA tensor loads
    assert "LDTM" in sass         # tcgen05.ld (TMEM -> registers)


def test_product_has_no_cpu_fallback():
    import torch
    from cervix_b200 import backend
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    prev = backend.set_backend(None)
    try:
        with pytest.raises(_lib.CervixError):
            backend.get_backend()
    finally:
        backend.set_backend(prev)
