"""``f_score`` - the per-step training metric of the reference's ``utils/utils_metrics.py:13-35``
(the file/plot utilities of that module are outside the hot path)."""
import torch

from .. import ops
from ..nets.deeplabv3_training import _match_target_size


def f_score(inputs, target, beta=1, smooth=1e-5, threhold=0.5):
    """``target`` is the one-hot label tensor [N,H,W,C+1]; returns a 0-dim tensor (no grad)."""
    with torch.no_grad():
        inputs = _match_target_size(inputs, target.shape[1], target.shape[2])
        hard = target.argmax(-1)
        res = ops.seg_losses(inputs.float(), hard, target, None, 0.5, 2.0, beta, smooth, threhold)
        return res[3]
