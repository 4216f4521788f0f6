"""Drop-in for the reference's ``utils/dataloader.py``: re-exports the B200 implementation (augmentation on the device)."""
import _bootstrap  # noqa: F401
from cervix_b200.utils.dataloader import *  # noqa: F401,F403
from cervix_b200.utils import dataloader as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
