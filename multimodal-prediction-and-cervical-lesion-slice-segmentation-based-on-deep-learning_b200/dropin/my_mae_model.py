"""Drop-in for the reference's ``MM/*/my_mae_model.py``: re-exports the B200 fusion head."""
import _bootstrap  # noqa: F401
from cervix_b200.multimodal import my_mae_model as _impl
from cervix_b200.multimodal.my_mae_model import *  # noqa: F401,F403

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
