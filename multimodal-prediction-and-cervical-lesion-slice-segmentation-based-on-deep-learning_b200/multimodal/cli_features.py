"""Clinical ("cli") node features of the classifier graph: four 1024-wide rows per patient built from the age.

Reference: ``MultiModal Prediction/Graph_Structure(data_augmentation).py`` :70-127 -
  row 0  20-bin one-hot of ``age // 5`` tiled to 1024 (:81-87);
  row 1  the same function applied to the NORMALISED age in [-1, 1] (:98), whose ``int(x // 5)`` is 0 for x >= 0 and
         -1 (i.e. the LAST bin) for x < 0 - reproduced on purpose;
  row 2  ``nn.Embedding(max_age + 1, 1024)`` looked up at the raw age (:105-107);
  row 3  ``nn.Embedding(101, 1024)`` looked up at ``int((x + 1) / 2 * 100)`` (:113-115).
The two embedding tables are random and frozen in the reference (never trained, never saved separately); here they are
buffers of the module so that a checkpoint reproduces the features.  All rows are assembled on the device for a whole
batch of patients; the table look-ups run on the CUDA row-gather operator.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import rowops as R

NUM_CATEGORIES = 20
VECTOR_LENGTH = 1024


def age_to_one_hot(age, num_categories: int = NUM_CATEGORIES, vector_length: int = VECTOR_LENGTH) -> np.ndarray:
    """Host restatement of the reference helper (:81-87) for one age value."""
    category = int(age // 5)
    one_hot = np.zeros(num_categories)
    one_hot[category] = 1
    return np.tile(one_hot, vector_length // num_categories + 1)[:vector_length]


def normalize_ages(ages, age_min=None, age_max=None):
    """(age - mid) / range * 2 over the cohort (:72-74); min / max default to the batch's own."""
    a = torch.as_tensor(ages, dtype=torch.float64)
    lo = float(a.min()) if age_min is None else float(age_min)
    hi = float(a.max()) if age_max is None else float(age_max)
    return (a - (hi + lo) / 2) / (hi - lo) * 2


class AgeNodeFeatures(nn.Module):
    def __init__(self, max_age: int = 100, dim: int = VECTOR_LENGTH):
        super().__init__()
        self.dim = dim
        self.register_buffer("age_table", torch.randn(max_age + 1, dim))       # nn.Embedding init: N(0, 1)
        self.register_buffer("age_std_table", torch.randn(101, dim))

    def forward(self, ages, age_min=None, age_max=None) -> torch.Tensor:
        """ages: G integer ages -> fp32 ``[G, 4, dim]`` on the module's device (``x_cli`` of the patient graphs)."""
        dev = self.age_table.device
        a = torch.as_tensor(ages).to(torch.int64)
        norm = normalize_ages(a, age_min, age_max)
        cols = torch.arange(self.dim, device=dev) % NUM_CATEGORIES
        cat = torch.div(a, 5, rounding_mode="floor") % NUM_CATEGORIES                 # int(age // 5), negative wraps
        cat_std = torch.floor(norm / 5).to(torch.int64) % NUM_CATEGORIES
        row0 = (cols[None, :] == cat.to(dev)[:, None]).float()
        row1 = (cols[None, :] == cat_std.to(dev)[:, None]).float()
        idx_std = ((norm + 1) / 2 * 100).to(torch.int64).clamp_(0, 100)
        if dev.type == "cuda":
            row2 = R.RowsGather.apply(self.age_table, a.to(device=dev, dtype=torch.int32), None)
            row3 = R.RowsGather.apply(self.age_std_table, idx_std.to(device=dev, dtype=torch.int32), None)
        else:
            row2, row3 = self.age_table[a], self.age_std_table[idx_std]
        return torch.stack([row0, row1, row2, row3], dim=1)
