"""``fit_one_epoch`` - drop-in for the reference's ``utils/utils_fit.py:31-197`` (same signature, same side effects:
``loss_history.append_loss``, ``eval_callback.on_epoch_end``, the three checkpoint files and their names).

What is different underneath:
  * the objective is ONE fused statistics pass + ONE gradient pass over the logits (``seg_objective``) instead of
    three ATen loss chains and a fourth softmax for ``f_score`` (utils_fit.py:71-84);
  * running loss / f_score stay on the device and are read back only when the progress line is refreshed and at the
    end of the phase - the reference synchronises twice per step (``.item()``, utils_fit.py:123-124);
  * ``fp16=True`` selects the bf16 tensor-core engine of the drop-in ``DeepLab`` (no loss scaling is needed in bf16,
    so ``scaler`` may be None and is never stepped); ``fp16=False`` selects the fp32 exact-parity engine;
  * on a CUDA device with the drop-in ``DeepLab`` and the optimizers the reference builds (``optim.Adam`` /
    ``optim.SGD(nesterov)`` over ``model.parameters()``, one param group; train.py:472-476) the training phase runs on
    ``engine.SegTrainer``: flat parameters, ONE fused optimizer launch per run of trainable parameters with the
    hyper-parameters of ``optimizer.param_groups[0]`` re-read every epoch (``set_optimizer_lr``, train.py:575), the whole
    step replayed as a CUDA graph from the third batch on, host batches prefetched one step ahead, bucketed NCCL
    all-reduce when ``model_train`` is ``DistributedDataParallel`` (train.py:386).  The math is torch.optim's; the torch
    optimizer object itself is left untouched (its ``state`` stays empty) and the moments live in the trainer, saved next
    to the weights as ``last_epoch_trainer_state.pth`` and restored from there when a run resumes.  Frozen parameters
    (Freeze_Train, train.py:447-449,531-551) are skipped and join later with their own step count, as in torch.
    ``CERVIX_FIT_EAGER=1`` keeps the plain autograd + ``optimizer.step()`` loop.
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


def _to_device(batch, cls_weights, cuda, local_rank, num_classes=None):
    imgs, pngs, labels = batch
    weights = torch.from_numpy(cls_weights) if not torch.is_tensor(cls_weights) else cls_weights
    if cuda:
        dev = torch.device("cuda", local_rank)
        imgs, pngs, labels = (None if t is None else t.to(dev, non_blocking=True) for t in (imgs, pngs, labels))
        weights = weights.to(dev)
    if pngs.dtype == torch.uint8:       # batches of utils.dataloader.DeviceAugmentLoader: the loader tail of dataloader.py:40-47
        pngs = pngs.long().clamp_(max=num_classes)
    if labels is None:
        labels = torch.nn.functional.one_hot(pngs, num_classes + 1).float()
    return imgs, pngs, labels, weights


def _wrap_loader(gen, cuda, local_rank):
    """A DataLoader built with the drop-in's ``deeplab_dataset_collate`` yields packed ``AugmentPlan`` batches (decoded
    uint8 sources + drawn decisions): run them through the augmentation kernels on the training device."""
    from .dataloader import DeviceAugmentLoader, deeplab_dataset_collate
    if getattr(gen, "collate_fn", None) is not deeplab_dataset_collate:
        return gen
    if not (cuda and torch.cuda.is_available()):
        raise RuntimeError("cervix_b200: the device augmentation of utils.dataloader needs a CUDA device (no CPU fallback)")
    return DeviceAugmentLoader(gen, torch.device("cuda", local_rank))


_FIT_STREAMS = {}


def _fit_stream(device):
    """The whole fast training phase runs on ONE dedicated non-default stream per device: autograd's AccumulateGrad nodes
    remember the stream they were created on, and a node that lives on the legacy default stream cannot take part in a
    later CUDA-graph capture (cudaErrorStreamCaptureImplicit: "would make the legacy stream depend on a capturing
    blocking stream")."""
    key = (device.type, device.index)
    if key not in _FIT_STREAMS:
        _FIT_STREAMS[key] = torch.cuda.Stream(device)
    return _FIT_STREAMS[key]


_TRAINER_STATE = "last_epoch_trainer_state.pth"
_EAGER_STEPS_BEFORE_CAPTURE = 2     # lazy initialisation (kernel attributes, allocator, NCCL) happens in real steps


def _fast_trainer(model_train, optimizer, cuda, dice_loss, focal_loss, cls_weights, num_classes, save_dir, epoch):
    """The ``SegTrainer`` that stands in for ``optimizer`` on this model, or None when the fast path does not apply."""
    import torch.distributed as dist
    from ..engine import SegTrainer
    from ..nets.deeplabv3_plus import DeepLab
    if not cuda or os.environ.get("CERVIX_FIT_EAGER") == "1":
        return None
    net = _unwrap(model_train)
    if not isinstance(net, DeepLab) or not next(net.parameters()).is_cuda:
        return None
    ddp = isinstance(model_train, torch.nn.parallel.DistributedDataParallel)
    if model_train is not net and not ddp:          # nn.DataParallel (train.py:388): single-process replicas, eager path
        return None
    if len(optimizer.param_groups) != 1:
        return None
    g = optimizer.param_groups[0]
    if {id(p) for p in g["params"]} != {id(p) for p in net.parameters()}:
        return None
    if type(optimizer) is torch.optim.Adam:
        if g.get("amsgrad") or g.get("maximize"):
            return None
        kind = "adam"
        kw = dict(betas=tuple(g["betas"]), eps=g["eps"])
    elif type(optimizer) is torch.optim.SGD:
        if g.get("dampening", 0) != 0 or g.get("maximize"):
            return None
        kind = "sgd"
        kw = dict(momentum=g["momentum"], nesterov=bool(g["nesterov"]))
    else:
        return None
    world = dist.get_world_size() if (ddp and dist.is_available() and dist.is_initialized()) else 1
    key = (id(optimizer), kind, bool(dice_loss), bool(focal_loss), world, tuple(float(w) for w in cls_weights))
    tr = getattr(net, "_cvx_trainer", None)
    if tr is None or getattr(tr, "_fit_key", None) != key:
        if tr is not None:
            tr.close()          # another optimizer / objective took over: release the old trainer's hooks and graph
        tr = SegTrainer(net, lr=g["lr"], weight_decay=g["weight_decay"], optimizer=kind, cls_weights=cls_weights,
                        num_classes=num_classes, dice=bool(dice_loss), focal=bool(focal_loss), world_size=world, **kw)
        tr._fit_key = key
        path = os.path.join(save_dir, _TRAINER_STATE) if save_dir else None
        if epoch > 0 and path and os.path.exists(path):     # a resumed run (Init_Epoch > 0): continue from the saved moments
            try:
                tr.load_state_dict(torch.load(path, map_location="cpu"))
            except Exception as e:
                print("cervix_b200: optimizer state at %s not restored (%s)" % (path, e))
        net._cvx_trainer = tr
    tr.sync_hyper(g)
    return tr


def _implicit_onehot(pngs, labels, num_classes) -> bool:
    """True when ``labels`` is exactly ``eye(C+1)[png]`` (what DeeplabDataset builds, dataloader.py:47): the loss kernels
    then derive the one-hot target from the class map and the 6.3 MB per image tensor need not cross PCIe."""
    if labels is None or labels.dim() != 4 or labels.shape[-1] != num_classes + 1 or labels.shape[:3] != pngs.shape:
        return False
    lab = labels.reshape(-1, num_classes + 1)
    idx = pngs.reshape(-1).long().clamp(0, num_classes)
    return bool((lab.sum(1) == 1).all() and (lab.gather(1, idx[:, None]) == 1).all())


def _train_phase_fast(trainer, gen, epoch_step, cuda, local_rank, cls_weights, num_classes, dice_loss, focal_loss, bar, main,
                      optimizer):
    from ..engine import BatchPrefetcher
    dev = torch.device("cuda", local_rank)
    state = {"implicit": None}

    def host_batches():
        for iteration, batch in enumerate(gen):
            if iteration >= epoch_step:
                return
            imgs, pngs, labels = batch
            if state["implicit"] is None:       # checked once per epoch on the host, before anything is copied
                state["implicit"] = labels is None or ((not labels.is_cuda) and _implicit_onehot(pngs, labels, num_classes))
            yield (imgs, pngs, None if state["implicit"] else labels)

    # a loader that already delivers device tensors (utils.dataloader.DeviceAugmentLoader: uint8 pixels and class maps
    # augmented on the GPU, uploaded one batch ahead on its own copy stream) needs no second staging step
    batches = host_batches() if getattr(gen, "on_device", False) else BatchPrefetcher(host_batches(), dev)
    run = None
    steps = 0
    for imgs, pngs, labels in batches:
        if trainer.graph_matches(imgs, pngs, labels):
            out = trainer.step_graphed(imgs, pngs, labels)
        elif trainer.t >= _EAGER_STEPS_BEFORE_CAPTURE and steps >= _EAGER_STEPS_BEFORE_CAPTURE:
            trainer.capture(imgs, pngs, labels, warmup=0)       # recording does not execute: replay it for this batch
            out = trainer.step_graphed(imgs, pngs, labels)
        else:
            out = trainer.step(imgs, pngs, labels)
        with torch.no_grad():                   # (ce, focal, dice, f_score) stay on the device
            run = out.clone() if run is None else run + out
        steps += 1
        if main and (steps % _REFRESH == 0 or steps == epoch_step):
            r = run.tolist()
            loss = (r[1] if focal_loss else r[0]) + (r[2] if dice_loss else 0.0)
            bar.update(total_loss=loss / steps, f_score=r[3] / steps, lr=get_lr(optimizer))
        else:
            bar.update()
    if run is None:
        return 0.0
    r = run.tolist()
    return (r[1] if focal_loss else r[0]) + (r[2] if dice_loss else 0.0)


def fit_one_epoch(model_train, model, loss_history, eval_callback, optimizer, epoch, epoch_step, epoch_step_val, gen,
                  gen_val, Epoch, cuda, dice_loss, focal_loss, cls_weights, num_classes, fp16, scaler, save_period,
                  save_dir, local_rank=0):
    _select_engine(model_train, fp16)
    gen, gen_val = _wrap_loader(gen, cuda, local_rank), _wrap_loader(gen_val, cuda, local_rank)
    main = local_rank == 0
    if main:
        print("Start Train")
    bar = _Progress(epoch_step, f"Epoch {epoch + 1}/{Epoch}", main)
    model_train.train()
    run_loss = run_fs = None
    steps = 0
    trainer = None
    if cuda and torch.cuda.is_available():
        dev = torch.device("cuda", local_rank)
        side, cur = _fit_stream(dev), torch.cuda.current_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            trainer = _fast_trainer(model_train, optimizer, cuda, dice_loss, focal_loss, cls_weights, num_classes, save_dir,
                                    epoch)
            if trainer is not None:
                total_fast = _train_phase_fast(trainer, gen, epoch_step, cuda, local_rank, cls_weights, num_classes,
                                               dice_loss, focal_loss, bar, main, optimizer)
                gen = ()
        cur.wait_stream(side)
    for iteration, batch in enumerate(gen):
        if iteration >= epoch_step:
            break
        with torch.no_grad():
            imgs, pngs, labels, weights = _to_device(batch, cls_weights, cuda, local_rank, num_classes)
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
    if trainer is not None:
        total_loss = total_fast
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
            imgs, pngs, labels, weights = _to_device(batch, cls_weights, cuda, local_rank, num_classes)
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
    if trainer is not None:     # optimizer moments + step counts, which the reference's checkpoints omit
        torch.save(trainer.state_dict(), os.path.join(save_dir, _TRAINER_STATE))
