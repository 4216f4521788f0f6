"""Drop-in for the reference's ``MM/Two_Modal/my_mae_model_2_NA.py``: re-exports the B200 fusion head under the class name the
train scripts import (``from my_mae_model_2_NA import fusion_model_mae_2``)."""
import _bootstrap  # noqa: F401
from cervix_b200.multimodal import my_mae_model as _impl
from cervix_b200.multimodal.my_mae_model import *  # noqa: F401,F403

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
fusion_model_mae_2 = _impl.fusion_model_mae_two_NA
