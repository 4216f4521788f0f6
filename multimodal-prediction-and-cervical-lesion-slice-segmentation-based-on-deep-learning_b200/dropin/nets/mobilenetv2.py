"""Drop-in for the reference's ``nets/mobilenetv2.py``: re-exports the B200 implementation."""
import _bootstrap  # noqa: F401
from cervix_b200.nets.mobilenetv2 import *  # noqa: F401,F403
from cervix_b200.nets import mobilenetv2 as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
