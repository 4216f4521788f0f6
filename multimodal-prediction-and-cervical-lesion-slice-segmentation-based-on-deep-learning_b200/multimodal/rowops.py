"""Autograd shells of the fusion head's row operators (kernels: csrc/rowops.cu).  All tensors are fp32
``[rows, C]`` matrices; a batch of G patient graphs is stored as G consecutive row segments."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import ops
from ..backend import get_backend


def linear(x: torch.Tensor, lin: torch.nn.Linear) -> torch.Tensor:
    """nn.Linear on a [rows, C_in] matrix through the 1x1 convolution kernels."""
    rows = x.shape[0]
    w = lin.weight.reshape(lin.out_features, lin.in_features, 1, 1)
    y = ops.conv2d(x.reshape(rows, 1, 1, lin.in_features), w, lin.bias, 1, 0, 1)
    return y.reshape(rows, lin.out_features)


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
    masked tokens) * factor / B / 5.  logits: list of [B,4]; mae_out / mae_labels [B*T, C]; sel uint8 [B*T]."""

    @staticmethod
    def forward(ctx, labels, sel, weights, mse_weight, inv_count, mae_out, mae_labels, *logits):
        B = get_backend()
        loss = torch.zeros(1, dtype=torch.float32, device=mae_out.device)
        grads = [B.softmax_ce(l.contiguous(), labels, loss, w, True) for l, w in zip(logits, weights)]
        da, db = B.masked_mse(mae_out.contiguous(), mae_labels.contiguous(), sel, loss, mse_weight, inv_count, True)
        ctx.save_for_backward(da, db, *grads)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        da, db, *grads = ctx.saved_tensors
        return (None, None, None, None, None, da * g, db * g) + tuple(d * g for d in grads)
