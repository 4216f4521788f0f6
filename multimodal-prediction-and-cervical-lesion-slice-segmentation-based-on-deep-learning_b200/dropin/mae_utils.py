"""Drop-in for the reference's ``MM/*/mae_utils.py``: the names the train scripts and the model import from it."""
import _bootstrap  # noqa: F401
from cervix_b200.multimodal.my_mae_model import (Attention, Block, Mlp, generate_mask,  # noqa: F401
                                                 get_sinusoid_encoding_table)
