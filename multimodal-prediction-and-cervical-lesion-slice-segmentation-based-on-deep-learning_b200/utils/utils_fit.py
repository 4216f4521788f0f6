"""``fit_one_epoch`` - drop-in for the reference's ``utils/utils_fit.py:31-197`` (same signature, same side effects:
``loss_history.append_loss``, ``eval_callback.on_epoch_end``, the three checkpoint files and their names).

What is different underneath:
  * the objective is ONE fused statistics pass + ONE gradient pass over the logits (``seg_objective``) instead of
    three ATen loss chains and a fourth softmax for ``f_score`` (utils_fit.py:71-84);
  * running loss / f_score stay on the device and are read back only when the progress line is refreshed and at the
    end of the phase - the reference synchronises twice per step (``.item()``, utils_fit.py:123-124);
  * ``fp16=True`` selects the bf16 tensor-core engine of the drop-in ``DeepLab`` (no loss scaling is needed in bf16,
    so ``scaler`` may be None and is never stepped); ``fp16=False`` selects the fp32 exact-parity engine.
"""
import os

import torch

from ..nets.deeplabv3_training import seg_objective
from .utils import get_lr

try:  # tqdm is optional: a plain counter keeps the loop dependency-free
    from tqdm import tqdm
except Exception:  # pragma: no cover
    tqdm = None

_REFRESH = 10   # steps between device -> host reads of the running averages


class _Progress:
    def __init__(self, total, desc, enabled):
        self.bar = tqdm(total=total, desc=desc, postfix=dict, mininterval=0.3) if (enabled and tqdm is not None) else None

    def update(self, **postfix):
        if self.bar is not None:
            if postfix:
                self.bar.set_postfix(**postfix)
            self.bar.update(1)

    def close(self):
        if self.bar is not None:
            self.bar.close()


def _unwrap(model):
    return model.module if hasattr(model, "module") else model


def _select_engine(model, fp16):
    net = _unwrap(model)
    setter = getattr(net, "set_compute_dtype", None)
    if setter is not None and next(net.parameters()).is_cuda:
        setter(torch.bfloat16 if fp16 else torch.float32)


def _objective(outputs, pngs, labels, weights, num_classes, dice_loss, focal_loss):
    ce, focal, dice, fs = seg_objective(outputs, pngs, labels, weights, num_classes)
    loss = focal if focal_loss else ce
    if dice_loss:
        loss = loss + dice
    return loss, fs


def _to_device(batch, cls_weights, cuda, local_rank):
    imgs, pngs, labels = batch
    weights = torch.from_numpy(cls_weights) if not torch.is_tensor(cls_weights) else cls_weights
    if cuda:
        dev = torch.device("cuda", local_rank)
        imgs, pngs, labels = (t.to(dev, non_blocking=True) for t in (imgs, pngs, labels))
        weights = weights.to(dev)
    return imgs, pngs, labels, weights


def fit_one_epoch(model_train, model, loss_history, eval_callback, optimizer, epoch, epoch_step, epoch_step_val, gen,
                  gen_val, Epoch, cuda, dice_loss, focal_loss, cls_weights, num_classes, fp16, scaler, save_period,
                  save_dir, local_rank=0):
    _select_engine(model_train, fp16)
    main = local_rank == 0
    if main:
        print("Start Train")
    bar = _Progress(epoch_step, f"Epoch {epoch + 1}/{Epoch}", main)
    model_train.train()
    run_loss = run_fs = None
    steps = 0
    for iteration, batch in enumerate(gen):
        if iteration >= epoch_step:
            break
        with torch.no_grad():
            imgs, pngs, labels, weights = _to_device(batch, cls_weights, cuda, local_rank)
        optimizer.zero_grad()
        outputs = model_train(imgs)
        loss, fs = _objective(outputs, pngs, labels, weights, num_classes, dice_loss, focal_loss)
        loss.backward()
        optimizer.step()
        with torch.no_grad():
            run_loss = loss.detach().clone() if run_loss is None else run_loss + loss.detach()
            run_fs = fs.detach().clone() if run_fs is None else run_fs + fs.detach()
        steps += 1
        if main and (steps % _REFRESH == 0 or steps == epoch_step):
            bar.update(total_loss=float(run_loss) / steps, f_score=float(run_fs) / steps, lr=get_lr(optimizer))
        else:
            bar.update()
    total_loss = float(run_loss) if run_loss is not None else 0.0
    bar.close()

    if main:
        print("Finish Train")
        print("Start Validation")
    bar = _Progress(epoch_step_val, f"Epoch {epoch + 1}/{Epoch}", main)
    model_train.eval()
    run_loss = run_fs = None
    steps = 0
    for iteration, batch in enumerate(gen_val):
        if iteration >= epoch_step_val:
            break
        with torch.no_grad():
            imgs, pngs, labels, weights = _to_device(batch, cls_weights, cuda, local_rank)
            outputs = model_train(imgs)
            loss, fs = _objective(outputs, pngs, labels, weights, num_classes, dice_loss, focal_loss)
            run_loss = loss.detach().clone() if run_loss is None else run_loss + loss.detach()
            run_fs = fs.detach().clone() if run_fs is None else run_fs + fs.detach()
        steps += 1
        if main and (steps % _REFRESH == 0 or steps == epoch_step_val):
            bar.update(val_loss=float(run_loss) / steps, f_score=float(run_fs) / steps, lr=get_lr(optimizer))
        else:
            bar.update()
    val_loss = float(run_loss) if run_loss is not None else 0.0
    bar.close()

    if not main:
        return
    print("Finish Validation")
    mean_train, mean_val = total_loss / epoch_step, val_loss / epoch_step_val
    loss_history.append_loss(epoch + 1, mean_train, mean_val)
    eval_callback.on_epoch_end(epoch + 1, model_train)
    print("Epoch:" + str(epoch + 1) + "/" + str(Epoch))
    print("Total Loss: %.3f || Val Loss: %.3f " % (mean_train, mean_val))
    state = model.state_dict()
    if (epoch + 1) % save_period == 0 or epoch + 1 == Epoch:
        torch.save(state, os.path.join(save_dir, "ep%03d-loss%.3f-val_loss%.3f.pth" % (epoch + 1, mean_train, mean_val)))
    if len(loss_history.val_loss) <= 1 or mean_val <= min(loss_history.val_loss):
        print("Save best model to best_epoch_weights.pth")
        torch.save(state, os.path.join(save_dir, "best_epoch_weights.pth"))
    torch.save(state, os.path.join(save_dir, "last_epoch_weights.pth"))
