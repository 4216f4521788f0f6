"""Drop-in for the reference's ``deeplab.py``: re-exports the B200 predictor."""
import _bootstrap  # noqa: F401
from cervix_b200.deeplab import DeeplabV3  # noqa: F401
