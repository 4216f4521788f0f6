"""Metrics of the reference's ``utils/utils_metrics.py``: ``f_score`` (:13-35, the per-step training metric) and the
confusion-matrix family behind mIoU / mPA / accuracy (``fast_hist`` :37-47, ``per_class_iu`` :63, ``per_class_PA_Recall``
:85, ``per_class_Precision`` :107, ``per_Accuracy`` :117).  ``fast_hist`` runs on the device and can accumulate a whole
validation set into one matrix; the PNG-directory driver ``compute_mIoU`` and the plots are outside the hot path."""
import numpy as np
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


def fast_hist(a, b, n, hist=None):
    """Confusion matrix ``hist[gt][pred]`` of label map ``a`` against prediction ``b`` (numpy arrays or tensors of class
    indices; labels >= n, e.g. 255, are ignored).  Returns a numpy int64 [n, n] array like the reference; pass a device
    tensor as ``hist`` to accumulate without leaving the GPU (it is then returned as is)."""
    from ..backend import get_backend
    dev = hist.device if torch.is_tensor(hist) else (a.device if torch.is_tensor(a) and a.is_cuda else
                                                     torch.device("cuda" if torch.cuda.is_available() else "cpu"))
    ta = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).reshape(-1)
    tb = torch.as_tensor(np.ascontiguousarray(b) if isinstance(b, np.ndarray) else b).reshape(-1)
    if ta.numel() != tb.numel():
        raise ValueError("fast_hist: label and prediction sizes differ (%d vs %d)" % (ta.numel(), tb.numel()))
    ta = ta.clamp(0, 255).to(device=dev, dtype=torch.uint8) if ta.dtype != torch.uint8 else ta.to(dev)
    tb = tb.clamp(0, 255).to(device=dev, dtype=torch.uint8) if tb.dtype != torch.uint8 else tb.to(dev)
    out = get_backend().confusion_matrix(tb.contiguous(), ta.contiguous(), int(n), hist if torch.is_tensor(hist) else None)
    return out if torch.is_tensor(hist) else out.cpu().numpy()


def _np(hist):
    return hist.detach().cpu().numpy() if torch.is_tensor(hist) else np.asarray(hist)


def per_class_iu(hist):
    h = _np(hist)
    return np.diag(h) / np.maximum(h.sum(1) + h.sum(0) - np.diag(h), 1)


def per_class_PA_Recall(hist):
    h = _np(hist)
    return np.diag(h) / np.maximum(h.sum(1), 1)


def per_class_Precision(hist):
    h = _np(hist)
    return np.diag(h) / np.maximum(h.sum(0), 1)


def per_Accuracy(hist):
    h = _np(hist)
    return np.sum(np.diag(h)) / np.maximum(np.sum(h), 1)
