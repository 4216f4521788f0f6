"""Makes ``import cervix_b200`` work from the drop-in shims: the repo root (the parent of the product package
directory) must be importable."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
