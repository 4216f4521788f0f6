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


def test_mask_plan_tables_follow_the_reference_indexing():
    """MaskPlan: x[~mask] keeps the visible tokens in modality order (my_mae_model.py:143), the decoder sees visible
    tokens first then mask tokens, each with its own position row, and the un-shuffle restores modality order
    (:325-335); update() rewrites the same device tensors."""
    import numpy as np
    from cervix_b200.multimodal.my_mae_model import MaskPlan
    masks = np.array([[True, False, True, True], [True, True, True, False], [False, True, True, True]])
    plan = MaskPlan(masks, torch.device("cpu"))
    assert plan.n_vis == 1 and (plan.G, plan.T) == (3, 4)
    assert plan.vis_idx.tolist() == [1, 7, 8]                                   # g*T + visible modality
    assert plan.dec_idx.tolist() == [0, -1, -1, -1, 1, -1, -1, -1, 2, -1, -1, -1]
    assert plan.dec_pos.tolist() == [1, 0, 2, 3, 3, 0, 1, 2, 0, 1, 2, 3]        # visible first, then the masked ones
    # un-shuffle: token t of patient g sits at decoder slot (position of t in visible+masked order)
    assert plan.unshuffle.tolist() == [1, 0, 2, 3, 5, 6, 7, 4, 8, 9, 10, 11]
    assert plan.sel.tolist() == masks.reshape(-1).astype(int).tolist()
    ptrs = (plan.tables.data_ptr(), plan.sel.data_ptr())
    masks2 = np.roll(masks, 1, axis=1)
    plan.update(masks2)
    fresh = MaskPlan(masks2, torch.device("cpu"))
    assert (plan.tables.data_ptr(), plan.sel.data_ptr()) == ptrs
    assert torch.equal(plan.tables, fresh.tables) and torch.equal(plan.sel, fresh.sel)
    two_visible = np.array([[False, False, True, True], [True, False, False, True], [True, True, False, False]])
    with pytest.raises(ValueError):
        plan.update(two_visible)                                                 # visible count is part of the plan
    p2 = MaskPlan(two_visible, torch.device("cpu"))
    assert p2.n_vis == 2 and p2.vis_idx.tolist() == [0, 1, 5, 6, 10, 11]
