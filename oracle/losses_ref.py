"""CPU oracle for the segmentation objective (TEST INFRASTRUCTURE ONLY - see
oracle/deeplab_ref.py for the import rules and the parity pin).

Restates, from /root/reference/Segmentation/deeplabv3+/:
  * nets/deeplabv3_training.py:9-19   CE_Loss    -> ce_loss
  * nets/deeplabv3_training.py:21-36  Focal_Loss -> focal_loss
  * nets/deeplabv3_training.py:38-56  Dice_loss  -> dice_loss
  * utils/utils_metrics.py:13-35      f_score    -> f_score
  * nets/deeplabv3_training.py:81-117 get_lr_scheduler -> lr_at
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _match_size(logits, h, w):
    # the reference resizes only when BOTH dims differ (`h != ht and w != wt`)
    if logits.shape[2] != h and logits.shape[3] != w:
        logits = F.interpolate(logits, size=(h, w), mode="bilinear", align_corners=True)
    return logits


def _per_pixel_weighted_nll(logits, target, cls_weights, num_classes):
    """-w[t] * log_softmax(logits)[t], 0 where t == ignore_index.  Shapes: logits
    [N,C,H,W], target [N,H,W] -> [N*H*W]."""
    n, c, h, w = logits.shape
    flat = logits.permute(0, 2, 3, 1).reshape(-1, c)
    t = target.reshape(-1)
    logp = F.log_softmax(flat, dim=-1)
    valid = t != num_classes
    ts = torch.where(valid, t, torch.zeros_like(t))
    picked = logp.gather(1, ts[:, None])[:, 0]
    wt = cls_weights.to(logits.dtype)[ts]
    nll = torch.where(valid, -wt * picked, torch.zeros_like(picked))
    wsum = torch.where(valid, wt, torch.zeros_like(wt)).sum()
    return nll, wsum


def ce_loss(logits, target, cls_weights, num_classes=5):
    logits = _match_size(logits, target.shape[1], target.shape[2])
    nll, wsum = _per_pixel_weighted_nll(logits, target, cls_weights, num_classes)
    return nll.sum() / wsum  # weighted mean over non-ignored pixels


def focal_loss(logits, target, cls_weights, num_classes=5, alpha=0.5, gamma=2):
    logits = _match_size(logits, target.shape[1], target.shape[2])
    nll, _ = _per_pixel_weighted_nll(logits, target, cls_weights, num_classes)
    logpt = -nll                      # note: class-weighted log-prob, 0 on ignored pixels
    pt = torch.exp(logpt)
    if alpha is not None:
        logpt = logpt * alpha
    loss = -((1 - pt) ** gamma) * logpt
    return loss.mean()                # mean over ALL pixels, ignored ones included


def _dice_terms(prob, onehot):
    tp = (onehot[..., :-1] * prob).sum(dim=(0, 1))
    fp = prob.sum(dim=(0, 1)) - tp
    fn = onehot[..., :-1].sum(dim=(0, 1)) - tp
    return tp, fp, fn


def dice_loss(logits, onehot, beta=1, smooth=1e-5):
    n, c = logits.shape[:2]
    logits = _match_size(logits, onehot.shape[1], onehot.shape[2])
    prob = torch.softmax(logits.permute(0, 2, 3, 1).reshape(n, -1, c), -1)
    tgt = onehot.reshape(n, -1, onehot.shape[-1])
    tp, fp, fn = _dice_terms(prob, tgt)
    score = ((1 + beta ** 2) * tp + smooth) / ((1 + beta ** 2) * tp + beta ** 2 * fn + fp + smooth)
    return 1 - score.mean()


def f_score(logits, onehot, beta=1, smooth=1e-5, threshold=0.5):
    n, c = logits.shape[:2]
    logits = _match_size(logits, onehot.shape[1], onehot.shape[2])
    prob = torch.softmax(logits.permute(0, 2, 3, 1).reshape(n, -1, c), -1)
    hard = (prob > threshold).to(logits.dtype)
    tgt = onehot.reshape(n, -1, onehot.shape[-1])
    tp, fp, fn = _dice_terms(hard, tgt)
    score = ((1 + beta ** 2) * tp + smooth) / ((1 + beta ** 2) * tp + beta ** 2 * fn + fp + smooth)
    return score.mean()


def lr_at(lr_decay_type, lr, min_lr, total_iters, it, warmup_iters_ratio=0.1, warmup_lr_ratio=0.1,
          no_aug_iter_ratio=0.3, step_num=10):
    """Learning rate at iteration ``it`` (deeplabv3_training.py:81-117)."""
    if lr_decay_type == "cos":
        warm = min(max(warmup_iters_ratio * total_iters, 1), 3)
        warm_start = max(warmup_lr_ratio * lr, 1e-6)
        no_aug = min(max(no_aug_iter_ratio * total_iters, 1), 15)
        if it <= warm:
            return (lr - warm_start) * (it / float(warm)) ** 2 + warm_start
        if it >= total_iters - no_aug:
            return min_lr
        return min_lr + 0.5 * (lr - min_lr) * (
            1.0 + math.cos(math.pi * (it - warm) / (total_iters - warm - no_aug)))
    decay = (min_lr / lr) ** (1 / (step_num - 1))
    step_size = total_iters / step_num
    if step_size < 1:
        raise ValueError("step_size must above 1.")
    return lr * decay ** (it // step_size)
