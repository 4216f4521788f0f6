"""Drop-in for the reference's ``utils/utils_fit.py``: re-exports the B200 implementation."""
import _bootstrap  # noqa: F401
from cervix_b200.utils.utils_fit import *  # noqa: F401,F403
from cervix_b200.utils import utils_fit as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
