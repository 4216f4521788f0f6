"""Training-step engine for the DeepLabv3+ path: flat parameter/gradient storage, one fused
optimizer launch, and bucketed NCCL gradient all-reduce overlapped with backward.

Replaces, for callers that want the fast path, the reference's per-step sequence
``optimizer.zero_grad(); outputs = model_train(imgs); loss = focal + dice; loss.backward();
optimizer.step()`` (utils/utils_fit.py:60-121) and its ``DistributedDataParallel`` wrapper
(train.py:386).  Semantics kept: torch.optim.Adam / SGD(nesterov) math, mean-reduced gradients
across ranks, per-rank BatchNorm statistics (the reference's default ``sync_bn=False``).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from . import ops
from .backend import get_backend
from .nets.deeplabv3_training import seg_objective


class FlatParams:
    """Parameters of a module re-homed into ONE contiguous fp32 buffer (and their ``.grad`` into a second one), so the
    optimizer and the all-reduce see a single tensor.  ``trainable_only=False`` takes every parameter, frozen ones
    included, so that a later ``requires_grad = True`` (the reference's Freeze_Train schedule, train.py:531-551) needs no
    re-homing and keeps the optimizer state."""

    def __init__(self, module: torch.nn.Module, trainable_only: bool = True):
        self.params = [p for p in module.parameters() if p.requires_grad or not trainable_only]
        dev = self.params[0].device
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4  # keep every tensor 16-byte aligned
        self.numel = off
        self.data = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad_views = []
        for p, o in zip(self.params, self.offsets):
            self.data[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.data[o:o + p.numel()].view(p.shape)
            self.grad_views.append(self.grad[o:o + p.numel()].view(p.shape))
        self.attach_grad_views()

    def padded(self, i: int) -> int:
        return (self.params[i].numel() + 3) // 4 * 4

    def attach_grad_views(self):
        """``p.grad`` = a view of the flat gradient: autograd then accumulates in place (one add kernel per parameter)."""
        for p, g in zip(self.params, self.grad_views):
            if p.requires_grad and p.grad is not g:
                p.grad = g

    def detach_grads(self):
        """``p.grad = None``: autograd then just keeps the gradient tensor each backward produced (no kernel), and
        ``GradGather`` copies them into the flat buffer one bucket per launch."""
        for p in self.params:
            p.grad = None


class GradGather:
    """One-launch gather of the gradients of parameters ``idx`` into ``FlatParams.grad`` (C ABI: cvx_multi_gather)."""

    def __init__(self, flat: FlatParams, idx: Optional[List[int]] = None):
        B = get_backend()
        ch = B.multi_gather_chunk()
        dev = flat.data.device
        self.idx = list(range(len(flat.params))) if idx is None else list(idx)
        sizes = [flat.params[i].numel() for i in self.idx]
        tids, starts = [], []
        for k, n in enumerate(sizes):
            for st in range(0, n, ch):
                tids.append(k)
                starts.append(st)
        self.flat = flat
        self.chunk_tensor = torch.tensor(tids, dtype=torch.int32, device=dev)
        self.chunk_start = torch.tensor(starts, dtype=torch.int32, device=dev)
        self.dst_offsets = torch.tensor([flat.offsets[i] for i in self.idx], dtype=torch.int64, device=dev)
        self.sizes = torch.tensor(sizes, dtype=torch.int64, device=dev)

    def new_table(self) -> torch.Tensor:
        return torch.zeros(len(self.idx), dtype=torch.int64, device=self.flat.data.device)

    def pointers(self) -> List[int]:
        ps = self.flat.params
        return [0 if ps[i].grad is None else ps[i].grad.data_ptr() for i in self.idx]

    @staticmethod
    def pointers_of(grads) -> List[int]:
        return [0 if g is None else g.data_ptr() for g in grads]

    def launch(self, table: torch.Tensor):
        get_backend().multi_gather(table, self.chunk_tensor, self.chunk_start, self.dst_offsets, self.sizes, self.flat.grad)


class _Bucket:
    """A run of consecutive trainable parameters whose gradients are gathered and all-reduced together."""
    __slots__ = ("idx", "start", "end", "pending", "gather", "table_eager", "table_graph", "captured_ptrs", "wire")

    def __init__(self, idx, start, end):
        self.idx, self.start, self.end, self.pending = idx, start, end, 0
        self.gather = self.table_eager = self.table_graph = self.captured_ptrs = self.wire = None


class SegTrainer:
    """One data-parallel training step of DeepLab with the reference's objective (focal | CE, + dice; train.py:259-265) on
    the CUDA engine: what ``DistributedDataParallel`` (train.py:386) + ``optimizer.step()`` (utils_fit.py:60-121) do for
    the reference.

    * every parameter lives in one flat fp32 buffer; the trainable ones are updated by ONE fused Adam / SGD-nesterov
      launch per contiguous run (frozen parameters - the reference's Freeze_Train phase - are skipped exactly as
      ``torch.optim`` skips parameters without gradient; a parameter that is unfrozen later starts its own Adam step
      count, as it does in torch, and nobody's moments are reset);
    * gradients are produced by autograd as separate tensors and copied into the flat gradient one bucket per launch as
      soon as the bucket's last gradient exists (buckets are filled from the END of the parameter list: backward reaches
      the decoder first); with more than one rank each bucket's all-reduce starts right there and runs under the rest of
      backward; ranks start from rank 0's weights (broadcast at construction, as DDP does);
    * ``capture`` records the WHOLE step - forward, objective, backward, bucket gathers, NCCL all-reduces, optimizer -
      as one CUDA graph; learning rate, momentum / betas, weight decay and the step counts are read from device memory,
      so ``set_lr`` and the per-epoch schedule (train.py:575) act on replays."""

    def __init__(self, model: torch.nn.Module, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, optimizer: str = "adam", momentum: float = 0.9, nesterov: bool = True,
                 cls_weights=None, num_classes: int = 5, dice: bool = True, focal: bool = True,
                 bucket_mb: float = 32.0, world_size: int = 1, wire_dtype: Optional[torch.dtype] = None):
        if optimizer not in ("adam", "sgd"):
            raise ValueError("optimizer must be 'adam' or 'sgd' (train.py:472-476), got %r" % (optimizer,))
        self.model = model
        self.flat = FlatParams(model, trainable_only=False)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.optimizer, self.momentum, self.nesterov = optimizer, momentum, nesterov
        self.num_classes, self.dice, self.focal = num_classes, dice, focal
        dev = self.flat.data.device
        self.cls_weights = None if cls_weights is None else torch.as_tensor(cls_weights, dtype=torch.float32, device=dev)
        self.m = torch.zeros_like(self.flat.data)
        self.v = torch.zeros_like(self.flat.data) if optimizer == "adam" else None
        self.t = 0
        self.world = world_size
        self.wire_dtype = wire_dtype
        self.bucket_mb = bucket_mb
        self.last = None
        # device-side hyper-parameters + step counters: what a captured CUDA graph reads on every replay
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)          # steps taken (dropout streams)
        b1 = betas[0] if optimizer == "adam" else momentum
        self.hyper_dev = torch.tensor([lr, b1, betas[1], eps, weight_decay, 1.0 / max(world_size, 1)],
                                      dtype=torch.float32, device=dev)
        self.graph = None
        self._graph_key = None
        self._use_gather = self.flat.grad.is_cuda      # CPU tensors (host-logic tests): autograd accumulates into views
        self._nbt = [m.num_batches_tracked for m in model.modules()
                     if isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and m.num_batches_tracked is not None]
        self._t_start = {}                              # parameter index -> step count when it became trainable
        self._sig = None
        self.works = []
        if world_size > 1 and dist.is_available() and dist.is_initialized():
            # ranks start from rank 0's parameters and buffers, as DistributedDataParallel's constructor does
            dist.broadcast(self.flat.data, 0)
            for buf in model.buffers():
                dist.broadcast(buf.data, 0)
        self._hooked = set()                            # hooks can only be put on parameters that require grad
        self._hook_handles = []
        self._hooks_live = False
        self.refresh_trainable()

    # ------------------------------------------------------------------ trainable set, runs, buckets
    def refresh_trainable(self) -> bool:
        """Re-derive optimizer runs and gradient buckets from the parameters' ``requires_grad`` flags.  Called at every
        step (a tuple comparison); does work only when a flag changed (Freeze_Train -> unfreeze).  Optimizer state is
        kept; a captured graph is dropped (re-captured by the caller)."""
        sig = tuple(p.requires_grad for p in self.flat.params)
        if sig == self._sig:
            return False
        self._sig = sig
        flat, n = self.flat, len(self.flat.params)
        dev = flat.data.device
        for i in range(n):
            if sig[i]:
                self._t_start.setdefault(i, self.t)
                if i not in self._hooked:
                    self._hooked.add(i)
                    self._hook_handles.append(flat.params[i].register_post_accumulate_grad_hook(self._make_hook(i)))
            else:
                self._t_start.pop(i, None)
                flat.params[i].grad = None
        flat.grad.zero_()
        # optimizer runs: maximal stretches of consecutive trainable parameters with the same start step
        starts = sorted(set(self._t_start.values()))
        self._counter_of = {ts: k for k, ts in enumerate(starts)}
        self.step_devs = torch.tensor([self.t - ts for ts in starts] or [0], dtype=torch.int32, device=dev)
        self.runs = []                                  # [flat start, flat end, counter index]
        for i in range(n):
            if not sig[i]:
                continue
            k = self._counter_of[self._t_start[i]]
            a, b = flat.offsets[i], flat.offsets[i] + flat.padded(i)
            if self.runs and self.runs[-1][1] == a and self.runs[-1][2] == k:
                self.runs[-1][1] = b
            else:
                self.runs.append([a, b, k])
        # gradient buckets over the trainable parameters, from the end
        cap = int(self.bucket_mb * 1024 * 1024 / 4)
        self.buckets: List[_Bucket] = []
        self.bucket_of = [-1] * n
        cur, idx = 0, []
        trainable = [i for i in range(n) if sig[i]]
        for pos in range(len(trainable) - 1, -1, -1):
            i = trainable[pos]
            idx.append(i)
            cur += flat.padded(i)
            if cur >= cap or pos == 0:
                lo, hi = min(idx), max(idx)
                bk = _Bucket(sorted(idx), flat.offsets[lo], flat.offsets[hi] + flat.padded(hi))
                if self._use_gather:
                    bk.gather = GradGather(flat, bk.idx)
                    bk.table_eager, bk.table_graph = bk.gather.new_table(), bk.gather.new_table()
                for j in idx:
                    self.bucket_of[j] = len(self.buckets)
                self.buckets.append(bk)
                cur, idx = 0, []
        self.graph = None
        self._graph_key = None
        self._split = False
        self._graph_a = self._graph_b = self._split_keep = None
        return True

    def close(self):
        """Detach from the model: remove the gradient hooks (they hold the trainer - and its captured graph's memory pool -
        alive) and drop the graph.  The parameters keep living in the flat buffer."""
        for h in self._hook_handles:
            h.remove()
        self._hook_handles, self._hooked = [], set()
        self.graph = None
        self._graph_key = None
        self._split = False
        self._graph_a = self._graph_b = self._split_keep = None
        for bk in self.buckets:
            bk.gather = bk.table_eager = bk.table_graph = None

    # ------------------------------------------------------------------ data parallel
    def _make_hook(self, i):
        def hook(_p):
            if not self._hooks_live:
                return
            b = self.bucket_of[i]
            if b < 0:
                return
            bk = self.buckets[b]
            bk.pending += 1
            if bk.pending == len(bk.idx):
                self._bucket_ready(bk)
        return hook

    def _bucket_ready(self, bk: _Bucket):
        """All gradients of the bucket exist: copy them into the flat gradient (one launch) and start its all-reduce, which
        then runs on NCCL's stream under the rest of backward."""
        bk.pending = -(1 << 30)                         # fired
        if bk.gather is not None:
            ptrs = bk.gather.pointers()
            if torch.cuda.is_current_stream_capturing():
                bk.captured_ptrs = ptrs                 # the table is filled right after the capture ends
                bk.gather.launch(bk.table_graph)
            else:
                bk.table_eager.copy_(torch.tensor(ptrs, dtype=torch.int64))
                bk.gather.launch(bk.table_eager)
        if self.world > 1:
            seg = self.flat.grad[bk.start:bk.end]
            if self.wire_dtype is not None and self.wire_dtype != torch.float32:
                bk.wire = seg.to(self.wire_dtype)       # halves the bytes on NVLink; summation in the wire type
                self.works.append(dist.all_reduce(bk.wire, op=dist.ReduceOp.SUM, async_op=True))
            else:
                self.works.append(dist.all_reduce(seg, op=dist.ReduceOp.SUM, async_op=True))

    def _finish_buckets(self):
        for bk in self.buckets:                         # parameters that received no gradient never fire their hook
            if bk.pending >= 0:
                self._bucket_ready(bk)
            bk.pending = 0
        for w in self.works:
            w.wait()                                    # the compute stream waits for NCCL's stream (no host block on CUDA)
        self.works = []
        for bk in self.buckets:
            if bk.wire is not None:
                self.flat.grad[bk.start:bk.end].copy_(bk.wire)
                bk.wire = None

    # ------------------------------------------------------------------ step
    def step(self, imgs: torch.Tensor, pngs: torch.Tensor, labels: Optional[torch.Tensor] = None):
        """imgs: fp32 NCHW in [0,1] (or the decoded uint8 [B,H,W,3] pixels); pngs: int64 class map (num_classes =
        ignore; or the uint8 class map as read from the PNG, clamped on the device); labels: the
        reference's fp32 one-hot [B,H,W,C+1] (optional).  Returns the 4-vector
        (ce, focal, dice, f_score) as a device tensor (no host sync)."""
        self.refresh_trainable()
        losses = self._forward_backward(imgs, pngs, labels)
        self._update()
        self.t += 1
        return losses

    def _forward_backward(self, imgs, pngs, labels):
        if self._use_gather:
            self.flat.detach_grads()
        else:
            self.flat.attach_grad_views()
            self.flat.grad.zero_()
        if pngs.dtype == torch.uint8:      # loader tail on the device: png[png >= num_classes] = num_classes, as int64
            pngs = get_backend().finish_batch_u8(None, pngs.contiguous(), self.num_classes, torch.float32)[1]
        self.step_dev.add_(1)
        self.step_devs.add_(1)
        ops.set_step_counter(self.step_dev)
        B = get_backend()
        arena = getattr(B, "zero_arena_begin", None) if self.flat.data.is_cuda else None
        if arena is not None:       # one fill for all the step's fp64 reduction workspaces instead of ~360 memset nodes
            arena(self.flat.data.device)
        try:
            with ops.defer_batch_counters():
                out = self.model(imgs)
            if self.model.training and self._nbt:
                torch._foreach_add_(self._nbt, 1)
            ce, focal, dice, fs = seg_objective(out, pngs, labels, self.cls_weights, self.num_classes)
            loss = (focal if self.focal else ce) + (dice if self.dice else 0.0)
            for bk in self.buckets:
                bk.pending = 0
            self._hooks_live = True
            try:
                loss.backward()
            finally:
                self._hooks_live = False
        finally:
            if arena is not None:
                B.zero_arena_end()
        self._finish_buckets()
        self.last = torch.stack([ce.detach(), focal.detach(), dice.detach(), fs.detach()])
        return self.last

    def _update(self, lo: int = 0, hi: Optional[int] = None):
        """The optimizer on every run of trainable parameters (clipped to the flat range lo..hi); hyper-parameters and step
        counts come from device memory."""
        B = get_backend()
        d, g, m, v = self.flat.data, self.flat.grad, self.m, self.v
        hi = self.flat.numel if hi is None else hi
        for a, b, k in self.runs:
            a, b = max(a, lo), min(b, hi)
            if a >= b:
                continue
            cnt = self.step_devs[k:k + 1]
            if self.optimizer == "adam":
                B.adam_step_dev(d[a:b], g[a:b], m[a:b], v[a:b], self.hyper_dev, cnt)
            else:
                B.sgd_step_dev(d[a:b], g[a:b], m[a:b], self.hyper_dev, bool(self.nesterov))

    def set_lr(self, lr: float):
        if lr != self.lr:
            self.lr = lr
            self.hyper_dev[0:1].fill_(lr)

    def sync_hyper(self, group: dict):
        """Adopt lr / betas / momentum / weight decay of a ``torch.optim`` param group (the reference changes the learning
        rate every epoch through ``set_optimizer_lr``, train.py:575)."""
        self.set_lr(float(group["lr"]))
        wd = float(group.get("weight_decay", self.wd))
        if wd != self.wd:
            self.wd = wd
            self.hyper_dev[4:5].fill_(wd)

    # ------------------------------------------------------------------ two-phase backward (data parallel, CUDA graphs)
    _ENTRY_MODULES = ("conv1", "bn1", "conv2", "bn2", "block1", "block2", "block3")

    def _split_plan(self):
        """Parameter indices before / after the model's cut (``ops.cut_point`` at the end of Xception's entry flow), or None
        when the model has no cut, a parameter is frozen, or the early parameters are not one leading run of the flat
        buffer."""
        names = [n for n, _ in self.model.named_parameters()]
        if len(names) != len(self.flat.params) or not all(p.requires_grad for p in self.flat.params):
            return None
        early = [i for i, n in enumerate(names) if n.split(".")[0] == "backbone" and n.split(".")[1] in self._ENTRY_MODULES]
        if not early or early != list(range(len(early))) or type(getattr(self.model, "backbone", None)).__name__ != "Xception":
            return None
        late = list(range(len(early), len(names)))
        return early, late

    def _phase_a(self, imgs, pngs, labels, late):
        """Forward, objective, and the backward of everything AFTER the cut; gathers those gradients into the flat buffer."""
        if pngs.dtype == torch.uint8:
            pngs = get_backend().finish_batch_u8(None, pngs.contiguous(), self.num_classes, torch.float32)[1]
        self.step_dev.add_(1)
        self.step_devs.add_(1)
        ops.set_step_counter(self.step_dev)
        sink = {}
        ops._CUT_SINK[0] = sink
        try:
            with ops.defer_batch_counters():
                out = self.model(imgs)
        finally:
            ops._CUT_SINK[0] = None
        if self.model.training and self._nbt:
            torch._foreach_add_(self._nbt, 1)
        ce, focal, dice, fs = seg_objective(out, pngs, labels, self.cls_weights, self.num_classes)
        loss = (focal if self.focal else ce) + (dice if self.dice else 0.0)
        tags = sorted(sink)
        proxies = [sink[t][1] for t in tags]
        grads = torch.autograd.grad(loss, proxies + [self.flat.params[i] for i in late], allow_unused=True)
        self._split_cuts = ([sink[t][0] for t in tags], list(grads[:len(tags)]))
        self._split_late_grads = grads[len(tags):]          # kept alive until the gather of the captured graph is recorded
        self.last = torch.stack([ce.detach(), focal.detach(), dice.detach(), fs.detach()])
        return self.last

    def _phase_b(self, early):
        """Backward of the part BEFORE the cut from the gradients phase A left at the cut tensors."""
        outs, gouts = self._split_cuts
        keep = [(o, g) for o, g in zip(outs, gouts) if g is not None]
        grads = torch.autograd.grad([o for o, _ in keep], [self.flat.params[i] for i in early],
                                    grad_outputs=[g for _, g in keep], allow_unused=True)
        self._split_cuts = None
        return grads

    def capture_split(self, imgs: torch.Tensor, pngs: torch.Tensor, labels: Optional[torch.Tensor] = None, warmup: int = 3):
        """Data-parallel CUDA-graph step with communication under backward, without recording NCCL in a graph: the step is
        TWO graphs in one memory pool.  Graph A = forward + objective + backward down to the end of the entry flow + gather
        of those gradients (97 % of the parameters: middle flow, exit flow, ASPP, decoder); graph B = backward of the entry
        flow (3 % of the parameters, about a third of the backward time) + its gather.  Between the replays the host starts
        the all-reduce of A's gradients, which then runs on NCCL's stream under graph B; what is left exposed is the
        all-reduce of the entry flow's 6 MB.  Returns None (and captures nothing) when the model has no cut."""
        self.refresh_trainable()
        plan = self._split_plan()
        if plan is None or not self._use_gather:
            return None
        early, late = plan
        B = get_backend()
        dev = self.flat.data.device
        self.s_imgs, self.s_pngs = imgs.clone(), pngs.clone()
        self.s_labels = None if labels is None else labels.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(self.s_imgs, self.s_pngs, self.s_labels)
        torch.cuda.current_stream().wait_stream(side)
        flat = self.flat
        ga_late, ga_early = GradGather(flat, late), GradGather(flat, early)
        t_late, t_early = ga_late.new_table(), ga_early.new_table()
        self._split_ranges = ((flat.offsets[late[0]], flat.numel), (0, flat.offsets[late[0]]))
        pool = torch.cuda.graph_pool_handle()
        cap_stream = torch.cuda.Stream()
        self._graph_a, self._graph_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        arena = getattr(B, "zero_arena_begin", None)
        flat.detach_grads()
        try:
            with torch.cuda.graph(self._graph_a, pool=pool, stream=cap_stream):
                if arena is not None:
                    arena(dev)
                self.s_out = self._phase_a(self.s_imgs, self.s_pngs, self.s_labels, late)
                late_grads = self._split_late_grads
                ga_late.launch(t_late)
            with torch.cuda.graph(self._graph_b, pool=pool, stream=cap_stream):
                early_grads = self._phase_b(early)
                ga_early.launch(t_early)
        finally:
            if arena is not None:
                B.zero_arena_end()
        t_late.copy_(torch.tensor(GradGather.pointers_of(late_grads), dtype=torch.int64))
        t_early.copy_(torch.tensor(GradGather.pointers_of(early_grads), dtype=torch.int64))
        self._split_keep = (ga_late, ga_early, t_late, t_early, pool)
        self._split_late_grads = None
        self.graph = self._graph_a                       # "a captured step exists" for graph_matches / staleness checks
        self._graph_key = (tuple(imgs.shape), imgs.dtype, tuple(pngs.shape), pngs.dtype,
                           None if labels is None else tuple(labels.shape))
        self._comm_in_graph = False
        self._split = True
        return self

    def _step_split(self):
        (lo_a, hi_a), (lo_b, hi_b) = self._split_ranges
        g = self.flat.grad
        self._graph_a.replay()
        works = []
        if self.world > 1:
            works.append(dist.all_reduce(g[lo_a:hi_a], op=dist.ReduceOp.SUM, async_op=True))   # under graph B
        self._graph_b.replay()
        if self.world > 1:
            works.append(dist.all_reduce(g[lo_b:hi_b], op=dist.ReduceOp.SUM, async_op=True))
            works[0].wait()
        self._update(lo_a, hi_a)
        if self.world > 1:
            works[1].wait()
        self._update(lo_b, hi_b)

    # ------------------------------------------------------------------ checkpoint (optimizer + step state)
    def state_dict(self) -> dict:
        """Optimizer moments, step counts and hyper-parameters (the reference saves weights only, utils_fit.py:191-198; a
        run resumed from those restarts Adam from zero moments).  Weights themselves stay in ``model.state_dict()``."""
        n = len(self.flat.params)
        return {"optimizer": self.optimizer, "t": self.t, "m": self.m.detach().cpu(),
                "v": None if self.v is None else self.v.detach().cpu(),
                "t_start": [self._t_start.get(i, -1) for i in range(n)],
                "hyper": [float(x) for x in self.hyper_dev.cpu()], "numel": self.flat.numel,
                "step_dev": int(self.step_dev.item())}

    def load_state_dict(self, sd: dict):
        if sd["optimizer"] != self.optimizer or sd["numel"] != self.flat.numel:
            raise ValueError("trainer state is for optimizer %r with %d parameters" % (sd["optimizer"], sd["numel"]))
        self.m.copy_(sd["m"])
        if self.v is not None and sd["v"] is not None:
            self.v.copy_(sd["v"])
        self.t = int(sd["t"])
        self.step_dev.fill_(int(sd["step_dev"]))
        self._t_start = {i: ts for i, ts in enumerate(sd["t_start"]) if ts >= 0 and self.flat.params[i].requires_grad}
        self._sig = None
        self.refresh_trainable()
        return self

    # ------------------------------------------------------------------ CUDA graph
    def capture(self, imgs: torch.Tensor, pngs: torch.Tensor, labels: Optional[torch.Tensor] = None, warmup: int = 3,
                comm_in_graph: bool = False):
        """Capture one training step (~1 800 kernel launches) into a CUDA graph over static input buffers: forward,
        objective, backward, the bucket gathers and - single GPU - the optimizer.  Data parallel (default): the graph ends
        with the gathered flat gradient; after each replay the buckets are all-reduced back to back on NCCL's stream and
        the optimizer of bucket k runs as soon as ITS all-reduce is done, i.e. under the all-reduce of bucket k+1
        (``_reduce_update_pipelined``).  ``comm_in_graph=True`` records the bucket all-reduces inside the graph, forked
        where each bucket completes, so that they overlap backward as in the eager step; with torch 2.11 / NCCL 2.28 on the
        2 x B200 box that capture hung at replay (profiles/r02_multigpu_notes.txt), so it is opt-in."""
        self.refresh_trainable()
        self._split = False
        self._comm_in_graph = comm_in_graph or self.world == 1
        self.s_imgs, self.s_pngs = imgs.clone(), pngs.clone()
        self.s_labels = None if labels is None else labels.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(self.s_imgs, self.s_pngs, self.s_labels)
        torch.cuda.current_stream().wait_stream(side)
        world = self.world
        self.graph = torch.cuda.CUDAGraph()
        try:
            if not self._comm_in_graph:
                self.world = 1                          # no collective while capturing
            with torch.cuda.graph(self.graph):
                self.s_out = self._forward_backward(self.s_imgs, self.s_pngs, self.s_labels)
                if self._comm_in_graph:
                    self._update()
        finally:
            self.world = world
        for bk in self.buckets:
            if bk.captured_ptrs is not None:
                bk.table_graph.copy_(torch.tensor(bk.captured_ptrs, dtype=torch.int64))
                bk.captured_ptrs = None
        self._graph_key = (tuple(imgs.shape), imgs.dtype, tuple(pngs.shape), pngs.dtype,
                           None if labels is None else tuple(labels.shape))
        return self

    def graph_matches(self, imgs, pngs, labels=None) -> bool:
        return self.graph is not None and self._graph_key == (
            tuple(imgs.shape), imgs.dtype, tuple(pngs.shape), pngs.dtype, None if labels is None else tuple(labels.shape))

    def step_graphed(self, imgs: torch.Tensor, pngs: torch.Tensor, labels: Optional[torch.Tensor] = None):
        if self.refresh_trainable() or self.graph is None:
            raise RuntimeError("the captured step is stale (the trainable set changed): call capture() again")
        self.s_imgs.copy_(imgs, non_blocking=True)
        self.s_pngs.copy_(pngs, non_blocking=True)
        if self.s_labels is not None and labels is not None:
            self.s_labels.copy_(labels, non_blocking=True)
        if getattr(self, "_split", False):
            self._step_split()
        else:
            self.graph.replay()
            if not self._comm_in_graph:
                self._reduce_update_pipelined()
        self.t += 1
        return self.s_out

    def _reduce_update_pipelined(self, groups: int = 2):
        """All-reduce the flat gradient in `groups` contiguous pieces (neighbouring buckets merged: a 219 MB gradient as
        seven 32 MB calls costs more in per-call latency than it wins) on NCCL's stream; the compute stream waits for piece
        k only, then updates the parameters of piece k while piece k+1 is still on the wire."""
        n = len(self.buckets)
        per = (n + groups - 1) // groups
        pieces = []
        for g0 in range(0, n, per):                      # buckets run from the END of the flat buffer downwards
            bks = self.buckets[g0:g0 + per]
            pieces.append((min(b.start for b in bks), max(b.end for b in bks)))
        works, wires = [], []
        for lo, hi in pieces:
            seg = self.flat.grad[lo:hi]
            if self.wire_dtype is not None and self.wire_dtype != torch.float32:
                wire = seg.to(self.wire_dtype)
                wires.append(wire)
                works.append(dist.all_reduce(wire, op=dist.ReduceOp.SUM, async_op=True))
            else:
                wires.append(None)
                works.append(dist.all_reduce(seg, op=dist.ReduceOp.SUM, async_op=True))
        for (lo, hi), w, wire in zip(pieces, works, wires):
            w.wait()
            if wire is not None:
                self.flat.grad[lo:hi].copy_(wire)
            self._update(lo, hi)

class FusionTrainer:
    """One data-parallel training step of the multimodal fusion head over a batch of patients: forward of all
    patients at once (``fusion_model_mae_2.forward_batch``), the reference's objective (my_train(full).py:309-347:
    CE_all + 0.3 CE_img* + 0.2 CE_cli + the masked-auto-encoder MSE), backward, gradient sum-all-reduce across ranks
    (patients are sharded) and ``torch.optim.Adam(lr=1e-4, weight_decay=5e-4)`` math
    (my_train(full).py:233-236) as one fused launch on the flat parameters.  The reference steps once per 8
    patients; a rank's shard plays that role here and the all-reduced gradient is divided by the world size."""

    def __init__(self, head: torch.nn.Module, train_types, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 5e-4, world_size: int = 1):
        self.head = head
        self.train_types = list(train_types)
        self.flat = FlatParams(head)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.m = torch.zeros_like(self.flat.data)
        self.v = torch.zeros_like(self.flat.data)
        self.t = 0
        self.world = world_size
        # step count and hyper-parameters live in device memory for the eager step as well: the eager and the
        # graph-replayed step then run the SAME optimizer arithmetic (bias corrections in fp32 on the device), so their
        # parameters stay bit-identical - the head's gradients are ill-conditioned enough (LayerNorm over 32 post-ReLU
        # values) for a one-ulp difference in a weight to move some gradients by 1e-3
        dev = self.flat.data.device
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.hyper_dev = torch.tensor([lr, betas[0], betas[1], eps, weight_decay, 1.0 / max(world_size, 1)],
                                      dtype=torch.float32, device=dev)
        self._gather = GradGather(self.flat) if self.flat.grad.is_cuda else None
        self._table_eager = self._gather.new_table() if self._gather else None
        self._table_graph = self._gather.new_table() if self._gather else None
        self._captured_ptrs = None

    def _forward_backward(self, feats, edges, labels, masks, use_types=None, mix: bool = True):
        from .multimodal.my_mae_model import fusion_objective
        gather = self._gather is not None
        if gather:      # autograd keeps each parameter's gradient tensor (no accumulate kernel); ONE launch copies them all
            self.flat.detach_grads()
        else:
            self.flat.attach_grad_views()
            self.flat.grad.zero_()
        out = self.head.forward_batch(feats, edges, self.train_types, use_types or self.train_types, masks, mix)
        loss = fusion_objective(out, labels, masks)
        loss.backward()
        if gather:
            ptrs = self._gather.pointers()
            if torch.cuda.is_current_stream_capturing():
                self._captured_ptrs = ptrs              # the table is filled right after the capture ends
                self._gather.launch(self._table_graph)
            else:
                self._table_eager.copy_(torch.tensor(ptrs, dtype=torch.int64))
                self._gather.launch(self._table_eager)
        return loss.detach()

    def step(self, feats, edges, labels: torch.Tensor, masks, use_types=None, mix: bool = True) -> torch.Tensor:
        """feats[m]: ``[G, nodes_m, 1024]`` node features of this rank's G patients; masks: bool ``[G, T]`` (True =
        masked modality).  Returns this rank's loss (device scalar, no host sync)."""
        self.step_dev.add_(1)
        loss = self._forward_backward(feats, edges, labels, masks, use_types, mix)
        if self.world > 1:
            dist.all_reduce(self.flat.grad, op=dist.ReduceOp.SUM)
        self.t += 1
        get_backend().adam_step_dev(self.flat.data, self.flat.grad, self.m, self.v, self.hyper_dev, self.step_dev)
        return loss

    # ------------------------------------------------------------------ CUDA-graph step
    def capture(self, feats, edges, labels: torch.Tensor, masks, warmup: int = 2):
        """Capture the head's train step (~600 launches of a few microseconds each: the eager step is bound by the
        host, not the GPU) into a CUDA graph over static copies of the inputs.  The modality masks enter through a
        ``MaskPlan`` whose device-side index tables the graph re-reads, the dropout / attention-dropout streams and
        Adam's step count through a device counter, so every replay is a fresh training step.  Data parallel: forward
        + backward are captured; the all-reduce and the fused Adam follow each replay."""
        from .multimodal.my_mae_model import MaskPlan
        dev = self.flat.data.device
        self.s_feats = {m: t.detach().clone() for m, t in feats.items()}
        self.s_edges = edges
        self.s_labels = labels.clone()
        self.plan = MaskPlan(masks, dev)

        def body():
            self.step_dev.add_(1)
            ops.set_step_counter(self.step_dev)
            loss = self._forward_backward(self.s_feats, self.s_edges, self.s_labels, self.plan)
            if self.world == 1:
                get_backend().adam_step_dev(self.flat.data, self.flat.grad, self.m, self.v, self.hyper_dev, self.step_dev)
            return loss

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
                if self.world > 1:
                    self._reduce_and_update_dev()
                self.t += 1
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.s_loss = body()
        if self._captured_ptrs is not None:
            self._table_graph.copy_(torch.tensor(self._captured_ptrs, dtype=torch.int64))
            self._captured_ptrs = None
        # (the capture itself does not execute: step_dev still equals the number of steps taken)
        return self

    def _reduce_and_update_dev(self):
        dist.all_reduce(self.flat.grad, op=dist.ReduceOp.SUM)
        get_backend().adam_step_dev(self.flat.data, self.flat.grad, self.m, self.v, self.hyper_dev, self.step_dev)

    def step_graphed(self, feats, labels: torch.Tensor, masks) -> torch.Tensor:
        for m, t in feats.items():
            self.s_feats[m].copy_(t, non_blocking=True)
        self.s_labels.copy_(labels, non_blocking=True)
        self.plan.update(masks)
        self.graph.replay()
        if self.world > 1:
            self._reduce_and_update_dev()
        self.t += 1
        return self.s_loss


class BatchPrefetcher:
    """Moves pinned host batches to the device one step ahead on a side stream, so the
    host->device copy of step i+1 overlaps the compute of step i (replaces the synchronous
    ``imgs.cuda(local_rank)`` calls of utils/utils_fit.py:52-58).  The device side is two fixed buffer sets used
    alternately (no allocator traffic in the steady state): a set is overwritten only after the work that was
    enqueued on the compute stream while it was the current batch has finished (event recorded at the next
    ``__next__``), so a yielded batch stays valid until the one after next is requested."""

    def __init__(self, batches, device=None):
        self.it = iter(batches)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [None, None]
        self.consumed = [None, None]
        self.k = 0
        self.last_slot = None
        self.next = None
        self._fetch()

    def _fetch(self):
        try:
            host = next(self.it)
        except StopIteration:
            self.next = None
            return
        s = self.k & 1
        self.k += 1
        with torch.cuda.stream(self.stream):
            if self.consumed[s] is not None:
                self.stream.wait_event(self.consumed[s])
            bufs = self.slots[s]
            if bufs is None or len(bufs) != len(host) or any(
                    (b is None) != (t is None) or (b is not None and (b.shape != t.shape or b.dtype != t.dtype))
                    for b, t in zip(bufs, host)):
                bufs = tuple(None if t is None else torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host)
                for b in bufs:
                    if b is not None:
                        b.record_stream(torch.cuda.current_stream(self.device))
                self.slots[s] = bufs
            for b, t in zip(bufs, host):
                if b is not None:
                    b.copy_(t, non_blocking=True)
        self.next = (s, bufs)

    def __iter__(self):
        return self

    def __next__(self):
        if self.last_slot is not None:      # everything that used the previous batch has been enqueued by now
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.consumed[self.last_slot] = ev
        if self.next is None:
            raise StopIteration
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        s, cur = self.next
        self.last_slot = s
        self._fetch()
        return cur
