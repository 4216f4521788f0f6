"""Drop-in for the reference's ``MM/*/util.py`` (Logger, adjust_learning_rate, get_edge_index_full)."""
import _bootstrap  # noqa: F401
from cervix_b200.multimodal.util import Logger, adjust_learning_rate, get_edge_index_full  # noqa: F401
