"""CPU tests of the fusion head's host graph (cervix_b200.multimodal.my_mae_model) over the plain-torch emulation
of the C ABI, against the golden vectors of the reference's unmodified model."""
import pytest
import torch

import cervix_b200.backend as backend
from tests import fusion_cases as FC
from tests.emu_backend import EmuBackend


@pytest.fixture(autouse=True)
def emu():
    prev = backend.set_backend(EmuBackend())
    yield
    backend.set_backend(prev)


@pytest.mark.parametrize("tag", ["4modal", "3modal"])
def test_batched_forward_loss_grads_match_golden(tag):
    FC.check_batched(tag, torch.device("cpu"))


@pytest.mark.parametrize("tag", ["4modal", "3modal"])
def test_reference_signature_single_patient(tag):
    FC.check_single(tag, torch.device("cpu"))


def test_missing_modality_inference():
    FC.check_missing_modality(torch.device("cpu"))


@pytest.mark.parametrize("tag", FC.variant_tags())
def test_reference_variant_files(tag):
    """Two_Modal/my_mae_model_2*.py and Three_Modal/my_mae_model_three.py: schema, defaults, tuple layout, values."""
    FC.check_variant(tag, torch.device("cpu"))
