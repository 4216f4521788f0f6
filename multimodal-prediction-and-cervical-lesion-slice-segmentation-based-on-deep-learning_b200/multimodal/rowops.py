"""Autograd shells of the fusion head's row operators (kernels: csrc/rowops.cu).  All tensors are fp32
``[rows, C]`` matrices; a batch of G patient graphs is stored as G consecutive row segments."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import ops
from ..backend import get_backend


def linear_conv(x: torch.Tensor, lin: torch.nn.Linear) -> torch.Tensor:
    """nn.Linear on a [rows, C_in] matrix through the 1x1 convolution kernels (the r01 path; kept for A/B timing)."""
    rows = x.shape[0]
    w = lin.weight.reshape(lin.out_features, lin.in_features, 1, 1)
    y = ops.conv2d(x.reshape(rows, 1, 1, lin.in_features), w, lin.bias, 1, 0, 1)
    return y.reshape(rows, lin.out_features)


class GroupedLinear(Function):
    """A layer of independent ``nn.Linear`` calls - the per-modality branches of the fusion head
    (my_mae_model.py:404-416, 544, 550, 657-674, 706-769) - as ONE grouped GEMM launch (``cvx_gemm_grouped``), forward,
    data gradient and weight + bias gradient alike.  ``plan`` is a tuple of entries
    ``(in_idx, in_row0, rows, w_idx, b_idx | -1, out_idx, out_row0)``: rows ``in_row0 .. +rows`` of input tensor
    ``in_idx`` times weight ``w_idx`` go to rows ``out_row0 .. +rows`` of output ``out_idx`` - so the branches can read
    from and write into tensors that stack the modalities along the rows, which lets the row operators between the
    layers run once per stack instead of once per modality.  Weights stay in nn.Linear's [out, in] layout: no packing,
    no transposes.  ``tensors`` = inputs, then weights, then biases."""

    @staticmethod
    def forward(ctx, plan, n_in, n_w, out_shapes, *tensors):
        ins = [t.contiguous() for t in tensors[:n_in]]
        ws = list(tensors[n_in:n_in + n_w])
        bs = list(tensors[n_in + n_w:])
        outs = [torch.empty(shape, dtype=torch.float32, device=ins[0].device) for shape in out_shapes]
        probs = []
        for (ii, r0, rows, wi, bi, oi, o0) in plan:
            x, w = ins[ii], ws[wi].detach()
            n, k = w.shape
            assert x.shape[1] == k and outs[oi].shape[1] == n, "grouped linear: shape mismatch"
            probs.append(dict(a=x[r0:r0 + rows], b=w, c=outs[oi][o0:o0 + rows], bias=None if bi < 0 else bs[bi].detach(),
                              m=rows, n=n, k=k, lda_m=k, lda_k=1, ldb_n=k, ldb_k=1, ldc=n))
        get_backend().gemm_grouped(probs)
        ctx.save_for_backward(*ins, *ws)
        ctx.plan, ctx.n_in, ctx.n_w, ctx.n_b = plan, n_in, n_w, len(bs)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *dys):
        B = get_backend()
        saved = ctx.saved_tensors
        ins, ws = saved[:ctx.n_in], saved[ctx.n_in:]
        plan = ctx.plan
        dys = [None if d is None else d.contiguous() for d in dys]
        dev = ins[0].device
        # ---- data gradients: dX[rows, k] = dY[rows, n] W[n, k], written into one buffer per input tensor
        dxs = [None] * ctx.n_in
        covered = [0] * ctx.n_in
        probs = []
        for (ii, r0, rows, wi, bi, oi, o0) in plan:
            if not ctx.needs_input_grad[4 + ii] or dys[oi] is None:
                continue
            if dxs[ii] is None:
                dxs[ii] = torch.empty_like(ins[ii])
            n, k = ws[wi].shape
            covered[ii] += rows
            probs.append(dict(a=dys[oi][o0:o0 + rows], b=ws[wi].detach(), c=dxs[ii][r0:r0 + rows], m=rows, n=k, k=n,
                              lda_m=n, lda_k=1, ldb_n=1, ldb_k=k, ldc=k))
        for ii, dx in enumerate(dxs):
            if dx is not None and covered[ii] != dx.shape[0]:      # rows no entry touched have no gradient
                raise RuntimeError("grouped linear: the plan does not cover every row of input %d" % ii)
        if probs:
            B.gemm_grouped(probs)
        # ---- weight and bias gradients: dW[n, k] = dY^T[n, rows] X[rows, k]; db[n] = row sums of dY^T (same launch)
        dws = [None] * ctx.n_w
        dbs = [None] * ctx.n_b
        probs = []
        for (ii, r0, rows, wi, bi, oi, o0) in plan:
            if dys[oi] is None or not ctx.needs_input_grad[4 + ctx.n_in + wi]:
                continue
            if dws[wi] is not None:
                raise RuntimeError("grouped linear: a weight may appear in one plan entry only")
            n, k = ws[wi].shape
            dws[wi] = torch.empty((n, k), dtype=torch.float32, device=dev)
            db = None
            if bi >= 0 and ctx.needs_input_grad[4 + ctx.n_in + ctx.n_w + bi]:
                db = dbs[bi] = torch.empty((n,), dtype=torch.float32, device=dev)
            probs.append(dict(a=dys[oi][o0:o0 + rows], b=ins[ii][r0:r0 + rows], c=dws[wi], rowsum=db, m=n, n=k, k=rows,
                              lda_m=1, lda_k=n, ldb_n=1, ldb_k=k, ldc=k))
        if probs:
            B.gemm_grouped(probs)
        return (None, None, None, None, *dxs, *dws, *dbs)


def linear_group(entries):
    """``entries``: list of ``(x, row0, rows, lin, out_key, out_row0)`` - rows of tensor ``x`` through ``lin`` into rows of
    the output named ``out_key`` (outputs with the same key are one stacked tensor).  Returns {out_key: tensor}."""
    ins, in_ids, ws, bs, plan = [], {}, [], [], []
    out_rows, out_cols, out_keys = {}, {}, []
    for (x, r0, rows, lin, key, o0) in entries:
        if id(x) not in in_ids:
            in_ids[id(x)] = len(ins)
            ins.append(x)
        if key not in out_rows:
            out_rows[key], out_cols[key] = 0, lin.out_features
            out_keys.append(key)
        out_rows[key] = max(out_rows[key], o0 + rows)
        assert out_cols[key] == lin.out_features, "stacked outputs must have one width"
        ws.append(lin.weight)
        bi = -1
        if lin.bias is not None:
            bi = len(bs)
            bs.append(lin.bias)
        plan.append((in_ids[id(x)], int(r0), int(rows), len(ws) - 1, bi, out_keys.index(key), int(o0)))
    shapes = tuple((out_rows[k], out_cols[k]) for k in out_keys)
    outs = GroupedLinear.apply(tuple(plan), len(ins), len(ws), shapes, *ins, *ws, *bs)
    return dict(zip(out_keys, outs))


def linear(x: torch.Tensor, lin: torch.nn.Linear) -> torch.Tensor:
    """One nn.Linear on a [rows, C_in] matrix: a group of one (no weight packing, fixed-order sums)."""
    return linear_group([(x, 0, x.shape[0], lin, 0, 0)])[0]


class SegLayerNorm(Function):
    @staticmethod
    def forward(ctx, x, w, b, groups, seg, eps, mode):
        x = x.contiguous()
        y, stats = get_backend().seg_layernorm_fwd(x, w.detach().contiguous(), b.detach().contiguous(), groups, seg, eps, mode)
        ctx.save_for_backward(x, w, stats)
        ctx.cfg = (groups, seg, eps, mode)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, stats = ctx.saved_tensors
        groups, seg, eps, mode = ctx.cfg
        dx, dw, db = get_backend().seg_layernorm_bwd(dy.contiguous(), x, w.detach().contiguous(), stats, groups, seg, eps, mode)
        return dx, dw, db, None, None, None, None


def graph_layernorm(x, ln, groups, seg):
    """torch_geometric LayerNorm(mode='graph') over each group's seg x C block (eps added to the std)."""
    return SegLayerNorm.apply(x, ln.weight, ln.bias, groups, seg, ln.eps, 0)


def layernorm(x, ln: torch.nn.LayerNorm):
    return SegLayerNorm.apply(x, ln.weight, ln.bias, x.shape[0], 1, ln.eps, 1)


class SegTable:
    """Index tables of a stack of modality branches (csrc/rowops.cu "segment-table operators"): the node rows of the
    modalities are stacked modality-major ``[sum_m G * nodes_m, C]``; segment = one patient graph of one modality.
    ``order`` fixes the numbering of the segments and with it the row order of pooled outputs: ``"gm"`` patient-major
    (segment g * P + m: the token order of the masked auto-encoder), ``"mg"`` modality-major (segment m * G + g: the
    row blocks of the per-modality head GEMMs).  Built once per (G, node counts, order, device) and cached by the model."""

    def __init__(self, G: int, nodes, order: str, device):
        P = len(nodes)
        self.G, self.P, self.nodes, self.order = G, list(nodes), order, device
        base, off = [], 0
        for n in nodes:
            base.append(off)
            off += G * n
        self.rows = off
        self.row_start = base + [off]                       # host: rows of parameter set (modality) i
        self.segments = G * P
        self.max_len = max(nodes)
        start, length, sset = [0] * self.segments, [0] * self.segments, [0] * self.segments
        row_seg = [0] * off
        for m in range(P):
            for g in range(G):
                s = g * P + m if order == "gm" else m * G + g
                start[s], length[s], sset[s] = base[m] + g * nodes[m], nodes[m], m
                for i in range(nodes[m]):
                    row_seg[start[s] + i] = s
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=device)   # noqa: E731
        self.seg_start, self.seg_len, self.seg_set, self.row_seg = i32(start), i32(length), i32(sset), i32(row_seg)
        self._host_start, self._host_set = start, sset

    def seg_id(self, m: int, g: int) -> int:
        return g * self.P + m if self.order == "gm" else m * self.G + g

    def rows_of(self, m: int):
        return self.row_start[m], self.row_start[m + 1]


class SegTabLayerNorm(Function):
    """PyG graph-mode LayerNorm (mode 0) / nn.LayerNorm (mode 1, segments of one row) of every modality branch in one
    launch, each segment with its own modality's affine parameters.  ``params`` = the weights, then the biases."""

    @staticmethod
    def forward(ctx, x, tab: SegTable, eps, mode, *params):
        n = len(params) // 2
        ws, bs = [p.detach().contiguous() for p in params[:n]], [p.detach().contiguous() for p in params[n:]]
        x = x.contiguous()
        y, stats = get_backend().segtab_layernorm_fwd(x, ws, bs, tab, eps, mode)
        ctx.save_for_backward(x, stats, *params[:n])
        ctx.tab, ctx.cfg, ctx.n = tab, (eps, mode), n
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats, *ws = ctx.saved_tensors
        dx, dws, dbs = get_backend().segtab_layernorm_bwd(dy.contiguous(), x, [w.detach().contiguous() for w in ws], stats,
                                                          ctx.tab, *ctx.cfg)
        return (dx, None, None, None, *dws, *dbs)


def segtab_layernorm(x, tab: SegTable, lns, mode: int = 0):
    return SegTabLayerNorm.apply(x, tab, lns[0].eps, mode, *[ln.weight for ln in lns], *[ln.bias for ln in lns])


class SegTabGatePool(Function):
    """my_GlobalAttention of every modality branch in one launch: pooled row s = softmax-weighted sum of segment s."""

    @staticmethod
    def forward(ctx, x, gate, tab: SegTable):
        x = x.contiguous()
        pooled, att = get_backend().segtab_gate_pool_fwd(x, gate.contiguous().reshape(-1), tab)
        ctx.save_for_backward(x, att)
        ctx.tab = tab
        ctx.mark_non_differentiable(att)
        return pooled, att

    @staticmethod
    def backward(ctx, dpooled, _datt):
        x, att = ctx.saved_tensors
        dx, dgate = get_backend().segtab_gate_pool_bwd(dpooled.contiguous(), x, att, ctx.tab)
        return dx, dgate.reshape(-1, 1), None


class SegTabBcastAdd(Function):
    """x[rows of segment s] += t[tok_of_seg[s]] for every modality at once (my_mae_model.py:636-649)."""

    @staticmethod
    def forward(ctx, x, t, tab: SegTable, tok_of_seg, seg_of_tok):
        ctx.tab, ctx.seg_of_tok, ctx.tokens = tab, seg_of_tok, t.shape[0]
        return get_backend().segtab_bcast_add(x.contiguous(), t.contiguous(), tab, tok_of_seg)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        return dy, get_backend().segtab_bcast_add_bwd(dy, ctx.tab, ctx.seg_of_tok, ctx.tokens), None, None, None


class GELU(Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return get_backend().gelu_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return get_backend().gelu_bwd(dy.contiguous(), x)


def gelu(x):
    return GELU.apply(x)


class GraphTopology:
    """CSR of a fixed directed graph (edge_index[0] = source, [1] = target) for mean aggregation onto the
    targets, plus the transposed CSR that the backward pass needs."""

    def __init__(self, edge_index: torch.Tensor, nodes: int, device):
        src, dst = edge_index[0].tolist(), edge_index[1].tolist()
        deg = [0] * nodes
        for d in dst:
            deg[d] += 1
        self.nodes = nodes

        def csr(rows, cols, weight_of):
            order = sorted(range(len(rows)), key=lambda e: rows[e])
            rowptr, col, w = [0] * (nodes + 1), [], []
            for e in order:
                rowptr[rows[e] + 1] += 1
                col.append(cols[e]); w.append(weight_of(e))
            for i in range(nodes):
                rowptr[i + 1] += rowptr[i]
            return (torch.tensor(rowptr, dtype=torch.int32, device=device), torch.tensor(col, dtype=torch.int32, device=device),
                    torch.tensor(w, dtype=torch.float32, device=device))

        self.fwd = csr(dst, src, lambda e: 1.0 / max(deg[dst[e]], 1))     # out[target] += x[source] / deg(target)
        self.bwd = csr(src, dst, lambda e: 1.0 / max(deg[dst[e]], 1))     # dx[source] += dout[target] / deg(target)


class GraphMean(Function):
    @staticmethod
    def forward(ctx, x, topo: GraphTopology, groups):
        ctx.topo, ctx.groups = topo, groups
        return get_backend().graph_gather(x.contiguous(), groups, topo.nodes, *topo.fwd)

    @staticmethod
    def backward(ctx, dy):
        return get_backend().graph_gather(dy.contiguous(), ctx.groups, ctx.topo.nodes, *ctx.topo.bwd), None, None


class GatePool(Function):
    @staticmethod
    def forward(ctx, x, gate, groups, seg):
        x = x.contiguous()
        pooled, att = get_backend().gate_pool_fwd(x, gate.contiguous().reshape(-1), groups, seg)
        ctx.save_for_backward(x, att)
        ctx.cfg = (groups, seg)
        ctx.mark_non_differentiable(att)
        return pooled, att

    @staticmethod
    def backward(ctx, dpooled, _datt):
        x, att = ctx.saved_tensors
        dx, dgate = get_backend().gate_pool_bwd(dpooled.contiguous(), x, att, *ctx.cfg)
        return dx, dgate.reshape(-1, 1), None, None


class AttnSmall(Function):
    @staticmethod
    def forward(ctx, qkv, b, n, h, d, scale, drop_p, seed):
        qkv = qkv.contiguous()
        step_dev = ops._STEP_DEV[0] if drop_p > 0.0 else None      # graph replays: a new dropout mask every step
        out, probs = get_backend().attn_small_fwd(qkv, b, n, h, d, scale, drop_p, seed, step_dev)
        ctx.save_for_backward(qkv, probs)
        ctx.cfg = (b, n, h, d, scale, drop_p, seed, step_dev)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, probs = ctx.saved_tensors
        return (get_backend().attn_small_bwd(dout.contiguous(), qkv, probs, *ctx.cfg),) + (None,) * 7


class L2Norm(Function):
    @staticmethod
    def forward(ctx, x):
        y, norms = get_backend().l2norm_fwd(x.contiguous())
        ctx.save_for_backward(y, norms)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, norms = ctx.saved_tensors
        return get_backend().l2norm_bwd(dy.contiguous(), y, norms)


class RowsGather(Function):
    """y[i] = x[idx[i]] (idx >= 0) or ``fill`` (idx < 0)."""

    @staticmethod
    def forward(ctx, x, idx, fill):
        ctx.save_for_backward(idx)
        ctx.src_rows, ctx.has_fill = x.shape[0], fill is not None
        ctx.fill_shape = None if fill is None else fill.shape
        return get_backend().rows_gather(x.contiguous(), idx, None if fill is None else fill.detach().contiguous().reshape(-1))

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        dx, dfill = get_backend().rows_scatter_add(dy.contiguous(), idx, ctx.src_rows, ctx.has_fill)
        return dx, None, (None if dfill is None else dfill.reshape(ctx.fill_shape))


class FusionObjective(Function):
    """my_train(full).py:309-347 for one mini-batch: CE(all) + sum_m w_m CE(m) + MSE(mae_out, mae_labels on the
    masked tokens) * factor / B / 5.  logits_all [B,4]; logits_stack [P*B,4] = the P modality heads' logits stacked
    modality-major (rows m*B .. (m+1)*B belong to modality m, weight weights[1+m]); mae_out / mae_labels [B*T, C];
    sel uint8 [B*T]."""

    @staticmethod
    def forward(ctx, labels, sel, weights, mse_weight, inv_count, mae_out, mae_labels, logits_all, logits_stack):
        B = get_backend()
        loss = torch.zeros(1, dtype=torch.float32, device=mae_out.device)
        logits_all, logits_stack = logits_all.contiguous(), logits_stack.contiguous()
        G = logits_all.shape[0]
        d_all = B.softmax_ce(logits_all, labels, loss, weights[0], True)
        d_stack = torch.empty_like(logits_stack)
        for i, w in enumerate(weights[1:]):
            B.softmax_ce(logits_stack[i * G:(i + 1) * G], labels, loss, w, True, out=d_stack[i * G:(i + 1) * G])
        da, db = B.masked_mse(mae_out.contiguous(), mae_labels.contiguous(), sel, loss, mse_weight, inv_count, True)
        ctx.save_for_backward(da, db, d_all, d_stack)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        da, db, d_all, d_stack = ctx.saved_tensors
        return (None, None, None, None, None, da * g, db * g, d_all * g, d_stack * g)
