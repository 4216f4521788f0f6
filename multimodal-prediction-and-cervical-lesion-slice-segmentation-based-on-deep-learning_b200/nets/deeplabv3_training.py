"""Segmentation objective, init and LR schedule - drop-in for the reference's
``nets/deeplabv3_training.py`` (CE_Loss :9-19, Focal_Loss :21-36, Dice_loss :38-56,
weights_init :58-76, get_lr_scheduler :81-117, set_optimizer_lr :119-122).

The three losses are views of ONE fused CUDA statistics pass + ONE gradient pass
(``ops.SegLoss``); ``seg_objective`` returns all of them (plus the f_score metric of
utils/utils_metrics.py:13-35) from a single pass for callers that want the reference's
default ``focal + dice`` objective without re-reading the logits three times.
"""
import math
from functools import partial

import torch
import torch.nn as nn

from .. import ops


def _match_target_size(inputs, ht, wt):
    n, c, h, w = inputs.size()
    if h != ht and w != wt:  # same (and-)condition as the reference
        x = ops.to_nhwc(inputs, torch.float32)
        inputs = ops.upsample_to_nchw(x, ht, wt)
    return inputs


def _as_weights(cls_weights, device):
    if cls_weights is None:
        return None
    if not torch.is_tensor(cls_weights):
        cls_weights = torch.as_tensor(cls_weights)
    return cls_weights.to(device=device, dtype=torch.float32)


def seg_objective(inputs, target, onehot=None, cls_weights=None, num_classes=None, alpha=0.5, gamma=2, beta=1,
                  smooth=1e-5, threshold=0.5):
    """(ce, focal, dice, f_score) as 0-dim tensors from one pass.  ``target`` is the int64
    class map (value ``num_classes`` = ignore); ``onehot`` the reference's [N,H,W,C+1] labels
    (optional - derived from ``target`` when omitted)."""
    inputs = _match_target_size(inputs, target.shape[1], target.shape[2])
    c = inputs.shape[1]
    if num_classes is not None and num_classes != c:
        raise ValueError("ignore_index=%d must equal the number of logit channels %d" % (num_classes, c))
    res = ops.seg_losses(inputs.float(), target.long(), onehot, _as_weights(cls_weights, inputs.device), alpha, gamma,
                         beta, smooth, threshold)
    return res[0], res[1], res[2], res[3].detach()


def CE_Loss(inputs, target, cls_weights, num_classes=5):
    return seg_objective(inputs, target, None, cls_weights, num_classes)[0]


def Focal_Loss(inputs, target, cls_weights, num_classes=5, alpha=0.5, gamma=2):
    return seg_objective(inputs, target, None, cls_weights, num_classes, alpha=alpha, gamma=gamma)[1]


def Dice_loss(inputs, target, beta=1, smooth=1e-5):
    """``target`` is the one-hot label tensor [N,H,W,C+1] (last channel = ignore/background pad)."""
    inputs = _match_target_size(inputs, target.shape[1], target.shape[2])
    hard = target.argmax(-1)
    res = ops.seg_losses(inputs.float(), hard, target, None, 0.5, 2.0, beta, smooth, 0.5)
    return res[2]


def weights_init(net, init_type='normal', init_gain=0.02):
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, 'weight') and classname.find('Conv') != -1:
            if init_type == 'normal':
                torch.nn.init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == 'xavier':
                torch.nn.init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == 'kaiming':
                torch.nn.init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
            elif init_type == 'orthogonal':
                torch.nn.init.orthogonal_(m.weight.data, gain=init_gain)
            else:
                raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
        elif classname.find('BatchNorm2d') != -1:
            torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
            torch.nn.init.constant_(m.bias.data, 0.0)
    print('initialize network with %s type' % init_type)
    net.apply(init_func)


def _warm_cos_lr(lr, min_lr, total_iters, warmup_total_iters, warmup_lr_start, no_aug_iter, iters):
    if iters <= warmup_total_iters:
        return (lr - warmup_lr_start) * pow(iters / float(warmup_total_iters), 2) + warmup_lr_start
    if iters >= total_iters - no_aug_iter:
        return min_lr
    return min_lr + 0.5 * (lr - min_lr) * (
        1.0 + math.cos(math.pi * (iters - warmup_total_iters) / (total_iters - warmup_total_iters - no_aug_iter)))


def _step_lr(lr, decay_rate, step_size, iters):
    if step_size < 1:
        raise ValueError("step_size must above 1.")
    return lr * decay_rate ** (iters // step_size)


def get_lr_scheduler(lr_decay_type, lr, min_lr, total_iters, warmup_iters_ratio=0.1, warmup_lr_ratio=0.1,
                     no_aug_iter_ratio=0.3, step_num=10):
    if lr_decay_type == "cos":
        warmup_total_iters = min(max(warmup_iters_ratio * total_iters, 1), 3)
        warmup_lr_start = max(warmup_lr_ratio * lr, 1e-6)
        no_aug_iter = min(max(no_aug_iter_ratio * total_iters, 1), 15)
        return partial(_warm_cos_lr, lr, min_lr, total_iters, warmup_total_iters, warmup_lr_start, no_aug_iter)
    decay_rate = (min_lr / lr) ** (1 / (step_num - 1))
    step_size = total_iters / step_num
    return partial(_step_lr, lr, decay_rate, step_size)


def set_optimizer_lr(optimizer, lr_scheduler_func, epoch):
    lr = lr_scheduler_func(epoch)
    for param_group in optimizer.param_groups:
        param_group['lr'] = lr
