"""B200 path of the multimodal fusion classifier.

Drop-in for ``MultiModal Prediction/{Four,Three,Two}_Modal/my_mae_model*.py``: the class
``fusion_model_mae_2`` keeps the reference's constructor (my_mae_model.py:400), its 148-entry
``state_dict`` schema and the 9-tuple ``forward`` (my_mae_model.py:500-793), so the reference's train scripts
(my_train(full).py:79-81,246-248) run unchanged.  One parametrised implementation covers the 2-, 3- and
4-modal variants (they differ only in ``train_type_num`` and in which modality branches are visited).

The reference processes ONE patient graph per forward through hundreds of tiny ATen / PyG launches.  Here
``forward_batch`` lays G patients out as ``[G * nodes, C]`` row matrices and every graph-level operator
(SAGE mean aggregation, PyG graph-mode LayerNorm, gated attention pooling, the <= 4-token MAE attention,
MLP-Mixer) is one CUDA launch over the whole batch (csrc/rowops.cu); linear layers go through the 1x1
convolution kernels.  ``forward`` is the G = 1 case.  No torch compute op is on the path: the modules below
are parameter holders and are never called.

torch_geometric / torch_scatter / timm are not needed: their published algorithms are
restated by the kernels and pinned by tests/golden/fusion_*.npz.
"""
from __future__ import annotations

import collections
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from . import rowops as R

MODALITIES = ("imgN", "imgA", "imgL", "cli")
_EDGE_ATTR = {"imgN": "edge_index_imageN", "imgA": "edge_index_imageA", "imgL": "edge_index_imageL",
              "cli": "edge_index_cli"}


# ----------------------------------------------------------------------------- modality-mask index tables
class MaskPlan:
    """Device-side index tables of one batch's modality masks (``[G, T]`` bool, True = masked; every patient has the
    same number of visible modalities, as ``generate_mask`` guarantees): which rows the MAE encoder keeps
    (``x[~mask]``, my_mae_model.py:143), how the decoder input is assembled from visible tokens and mask tokens, the
    position row of every decoder token, the un-shuffle back to modality order (:325-335) and the 0/1 selector of the
    masked rows for the reconstruction loss (my_train(full).py:253).  The tensors keep their addresses, so a captured
    CUDA graph of the train step reads new masks after ``update``."""

    def __init__(self, masks, device):
        masks = np.asarray(masks, dtype=bool)
        self.G, self.T = masks.shape
        self.device = device
        self.n_vis = int((~masks[0]).sum())
        lists, sel = self._lists(masks)
        self.tables = torch.tensor(lists, dtype=torch.int32, device=device)        # [4, G*T] (vis_idx is padded)
        self.sel = torch.tensor(sel, dtype=torch.uint8, device=device)
        n = self.G * self.n_vis
        self.vis_idx, self.dec_idx = self.tables[0, :n], self.tables[1]
        self.dec_pos, self.unshuffle = self.tables[2], self.tables[3]

    def _lists(self, masks):
        G, T, n_vis = self.G, self.T, self.n_vis
        if masks.shape != (G, T) or any(int((~m).sum()) != n_vis for m in masks):
            raise ValueError("all patients of a batch must have the same number of visible modalities")
        vis_idx, dec_idx, dec_pos, unshuffle = [], [], [], []
        for g in range(G):
            vis = [t for t in range(T) if not masks[g, t]]
            msk = [t for t in range(T) if masks[g, t]]
            vis_idx += [g * T + t for t in vis]
            dec_idx += [g * n_vis + j for j in range(n_vis)] + [-1] * len(msk)
            dec_pos += vis + msk
            slot = {t: j for j, t in enumerate(vis + msk)}
            unshuffle += [g * T + slot[t] for t in range(T)]
        vis_idx += [0] * (G * T - len(vis_idx))
        return [vis_idx, dec_idx, dec_pos, unshuffle], masks.reshape(-1).astype(np.uint8).tolist()

    def update(self, masks):
        """New masks of the same shape and visible count into the same device tensors (two small H2D copies)."""
        lists, sel = self._lists(np.asarray(masks, dtype=bool))
        self.tables.copy_(torch.tensor(lists, dtype=torch.int32))
        self.sel.copy_(torch.tensor(sel, dtype=torch.uint8))
        return self


# ----------------------------------------------------------------------------- parameter holders
class SAGEConv(nn.Module):
    """torch_geometric.nn.SAGEConv(aggr='mean', root_weight=True): lin_l(mean_j x_j) + lin_r(x_i)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)


class LayerNorm(nn.Module):
    """torch_geometric.nn.LayerNorm(mode='graph') parameter holder (eps is added to the std)."""

    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(in_channels))
        self.bias = nn.Parameter(torch.zeros(in_channels))


class my_GlobalAttention(nn.Module):
    def __init__(self, gate_nn, nn=None):
        super().__init__()
        self.gate_nn = gate_nn
        self.nn = nn


def _xavier(module):
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class Attention(nn.Module):
    """mae_utils.py:58-102: head_dim = dim // heads (512 // 12 = 42 -> qkv 512 -> 1512, proj 504 -> 512)."""

    def __init__(self, dim, num_heads, attn_drop):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, self.head_dim * num_heads * 3, bias=False)
        self.proj = nn.Linear(self.head_dim * num_heads, dim)
        self.attn_drop = attn_drop


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, attn_drop=0.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = Attention(dim, num_heads, attn_drop)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class _Encoder(nn.Module):
    def __init__(self, dim, heads, attn_drop):
        super().__init__()
        self.patch_embed = nn.Linear(dim, dim)
        self.blocks = nn.ModuleList([Block(dim, heads, 4.0, attn_drop)])
        self.norm = nn.LayerNorm(dim)
        _xavier(self)


class _Decoder(nn.Module):
    def __init__(self, dim, heads, attn_drop):
        super().__init__()
        self.blocks = nn.ModuleList([Block(dim, heads, 4.0, attn_drop)])
        self.norm = nn.LayerNorm(dim)
        self.head = nn.Linear(dim, dim)
        _xavier(self)


def get_sinusoid_encoding_table(n_position, d_hid):
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.tensor(table, dtype=torch.float32).unsqueeze(0)


class PretrainVisionTransformer(nn.Module):
    """my_mae_model.py:216-335 with encoder_depth = decoder_depth = 1, 12 / 8 heads, attn_drop 0.3
    (drop_path resolves to 0 at depth 1, init_values = 0 -> no layer scale)."""

    def __init__(self, dim=512, train_type_num=4, attn_drop_rate=0.3):
        super().__init__()
        self.encoder = _Encoder(dim, 12, attn_drop_rate)
        self.decoder = _Decoder(dim, 8, attn_drop_rate)
        self.encoder_to_decoder = nn.Linear(dim, dim, bias=False)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = get_sinusoid_encoding_table(train_type_num, dim)      # plain tensor, as in the reference
        nn.init.trunc_normal_(self.mask_token, std=0.02, a=-0.02, b=0.02)


def Mix_mlp(dim1):
    return nn.Sequential(nn.Linear(dim1, dim1), nn.GELU(), nn.Linear(dim1, dim1))


class MixerBlock(nn.Module):
    def __init__(self, dim1, dim2):
        super().__init__()
        self.norm = LayerNorm(dim2)
        self.mix_mip_1 = Mix_mlp(dim1)
        self.mix_mip_2 = Mix_mlp(dim2)


def GNN_relu_Block(dim2, dropout=0.3):
    return nn.Sequential(nn.ReLU(), LayerNorm(dim2), nn.Dropout(p=dropout))


def _gate(dim):
    return nn.Sequential(nn.Linear(dim, dim // 4), nn.ReLU(), nn.Linear(dim // 4, 1))


# ----------------------------------------------------------------------------- the model
class fusion_model_mae_2(nn.Module):
    # What the reference's per-variant copies of this class change: the default number of
    # training modalities, the default of ``mix``, four never-used ``norm3_*`` LayerNorms in the state_dict (156
    # instead of 148 entries), and which per-modality logits ``forward`` returns.  The subclasses at the end of this
    # module set them; the arithmetic is the same.
    _DEFAULT_TYPES = 4
    _DEFAULT_MIX = True
    _NORM3 = False
    _RETURN_LOGITS = ("imgN", "imgA", "imgL", "cli")

    def __init__(self, in_feats, n_hidden, out_classes, dropout=0.3, train_type_num=None):
        super().__init__()
        if train_type_num is None:
            train_type_num = self._DEFAULT_TYPES
        C = out_classes
        for m in MODALITIES:
            setattr(self, m + "_gnn_2", SAGEConv(in_feats, C))
            setattr(self, m + "_relu_2", GNN_relu_Block(C))
        self.fc_cli_1 = nn.Linear(1024, C)      # present in the reference's state_dict, never used in forward
        self.fc_cli_2 = nn.Linear(C, C)
        for m in MODALITIES:
            setattr(self, "mpool_" + m, my_GlobalAttention(_gate(C)))
        for m in MODALITIES:
            setattr(self, "mpool_" + m + "_2", my_GlobalAttention(_gate(C)))
        self.mae = PretrainVisionTransformer(C, train_type_num)
        self.mix = MixerBlock(train_type_num, C)
        for m in MODALITIES:
            setattr(self, "lin1_" + m, nn.Linear(C, C // 4))
            setattr(self, "lin2_" + m, nn.Linear(C // 4, C // 16))
            setattr(self, "lin3_" + m, nn.Linear(C // 16, C // 64))
        for m in MODALITIES:
            setattr(self, "norm1_" + m, LayerNorm(C // 4))
        for m in MODALITIES:
            setattr(self, "norm2_" + m, LayerNorm(C // 16))
        if self._NORM3:                        # my_mae_model_2*.py:479-482, my_mae_model_three.py:485-488 (never used)
            for m in MODALITIES:
                setattr(self, "norm3_" + m, LayerNorm(C // 64))
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(p=dropout)
        self.classifier = nn.Linear(C // 64, 4)
        for m in MODALITIES:
            setattr(self, "classifier_" + m, nn.Linear(C // 64, 4))
        self.hidden = C
        self.train_type_num = train_type_num
        self._topo: Dict[tuple, R.GraphTopology] = {}
        self._const: Dict[tuple, torch.Tensor] = {}
        self._plans: "collections.OrderedDict[tuple, MaskPlan]" = collections.OrderedDict()

    # ------------------------------------------------------------------ helpers
    def _topology(self, edge_index, nodes: int, device) -> R.GraphTopology:
        ei = torch.as_tensor(edge_index).detach().cpu().to(torch.int64)
        key = (nodes, str(device), ei.numpy().tobytes())
        t = self._topo.get(key)
        if t is None:
            t = self._topo[key] = R.GraphTopology(ei, nodes, device)
        return t

    def _idx(self, values: Sequence[int], device) -> torch.Tensor:
        key = ("idx", str(device), tuple(values))
        t = self._const.get(key)
        if t is None:
            t = self._const[key] = torch.tensor(list(values), dtype=torch.int32, device=device)
        return t

    def _pos_rows(self, order: Sequence[int], device) -> torch.Tensor:
        """Sinusoid table rows in the given token order, as a [len(order), C] device constant."""
        key = ("pos", str(device), tuple(order))
        t = self._const.get(key)
        if t is None:
            t = self._const[key] = self.mae.pos_embed[0][list(order)].contiguous().to(device)
        return t

    def _drop(self, x, p):
        return ops.dropout(x, p, self.training)

    def _block(self, x, blk: Block, b: int, n: int):
        a = blk.attn
        h = R.linear(R.layernorm(x, blk.norm1), a.qkv)
        p = a.attn_drop if self.training else 0.0
        seed = 0
        if p > 0.0:
            ops._dropout_counter[0] += 1
            seed = (torch.initial_seed() * 1000003 + ops._dropout_counter[0]) & (2 ** 63 - 1)
        h = R.AttnSmall.apply(h, b, n, a.num_heads, a.head_dim, a.scale, p, seed).reshape(b * n, a.num_heads * a.head_dim)
        x = ops.Add.apply(x, R.linear(h, a.proj))
        h = R.linear(R.gelu(R.linear(R.layernorm(x, blk.norm2), blk.mlp.fc1)), blk.mlp.fc2)
        return ops.Add.apply(x, h)

    def _plan(self, masks: np.ndarray, device) -> MaskPlan:
        """MaskPlan of a numpy mask batch, from a small LRU cache (training draws new masks every step)."""
        masks = np.asarray(masks, dtype=bool)
        key = (str(device), masks.shape, masks.tobytes())
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = MaskPlan(masks, device)
            if len(self._plans) > 64:
                self._plans.popitem(last=False)
        else:
            self._plans.move_to_end(key)
        return plan

    def _mae(self, tokens, plan: MaskPlan):
        """tokens [G*T, C] (modality order); plan: the index tables of the batch's masks.  Returns the reconstructed
        tokens [G*T, C] in modality order (my_mae_model.py:305-335)."""
        G, T, n_vis = plan.G, plan.T, plan.n_vis
        dev = tokens.device
        mae = self.mae
        x = R.linear(tokens, mae.encoder.patch_embed)
        x = ops.Add.apply(x, self._pos_rows(list(range(T)) * G, dev))
        x = R.RowsGather.apply(x, plan.vis_idx, None)                                  # x[~mask]
        x = self._block(x, mae.encoder.blocks[0], G, n_vis)
        x = R.linear(R.layernorm(x, mae.encoder.norm), mae.encoder_to_decoder)
        # decoder input: visible tokens first, then mask tokens, each plus its own position row
        x = R.RowsGather.apply(x, plan.dec_idx, mae.mask_token if n_vis < T else None)
        pos = R.RowsGather.apply(self._pos_rows(list(range(T)), dev), plan.dec_pos, None)
        x = ops.Add.apply(x, pos)
        x = self._block(x, mae.decoder.blocks[0], G, T)
        x = R.linear(R.layernorm(x, mae.decoder.norm), mae.decoder.head)
        return R.RowsGather.apply(x, plan.unshuffle, None)

    def _mixer(self, x, G: int, T: int):
        """MixerBlock (my_mae_model.py:345-369) on [G*T, C]: the same graph-mode LayerNorm twice, token mixing on
        the per-patient transpose, channel mixing."""
        C = x.shape[1]
        mix = self.mix
        y = R.graph_layernorm(x, mix.norm, G, T)
        y = ops.ToNCHW.apply(y.reshape(G, T, 1, C)).reshape(G * C, T)                  # per-patient transpose
        y = R.linear(R.gelu(R.linear(y, mix.mix_mip_1[0])), mix.mix_mip_1[2])
        y = ops.ToNHWC.apply(y.reshape(G, C, T, 1), torch.float32).reshape(G * T, C)
        x = ops.Add.apply(x, y)
        y = R.graph_layernorm(x, mix.norm, G, T)
        y = R.linear(R.gelu(R.linear(y, mix.mix_mip_2[0])), mix.mix_mip_2[2])
        return ops.Add.apply(x, y)

    def _segtab(self, G: int, nodes: Sequence[int], order: str, device) -> R.SegTable:
        key = ("segtab", G, tuple(nodes), order, str(device))
        t = self._const.get(key)
        if t is None:
            t = self._const[key] = R.SegTable(G, list(nodes), order, device)
        return t

    @staticmethod
    def _per_set(x, tab: R.SegTable, lins: Sequence[nn.Linear]):
        """The same layer of every modality branch on the stack ``x`` (own weights, own row block): ONE grouped GEMM."""
        return R.linear_group([(x, tab.row_start[i], tab.row_start[i + 1] - tab.row_start[i], lin, 0, tab.row_start[i])
                               for i, lin in enumerate(lins)])[0]

    def _pool_stack(self, x, present: Sequence[str], suffix: str, tab: R.SegTable):
        """Gated attention pooling (my_GlobalAttention, my_mae_model.py:35-63) of every modality at once: each gate layer is
        one grouped GEMM over the modalities' row blocks, the per-graph softmax + weighted sum one launch over all
        segments; pooled rows come out in the table's segment order."""
        pools = [getattr(self, "mpool_" + m + suffix) for m in present]
        h = ops.relu(self._per_set(x, tab, [pl.gate_nn[0] for pl in pools]))
        gate = self._per_set(h, tab, [pl.gate_nn[2] for pl in pools])
        return R.SegTabGatePool.apply(x, gate, tab)

    # ------------------------------------------------------------------ batched forward
    def forward_batch(self, feats: Dict[str, torch.Tensor], edges: Dict[str, torch.Tensor],
                      train_use_type: Sequence[str], use_type: Optional[Sequence[str]] = None,
                      masks=None, mix: Optional[bool] = None) -> Dict[str, torch.Tensor]:
        """feats[m]: fp32 ``[G, nodes_m, in_feats]`` on the device; edges[m]: ``[2, E]`` topology shared by all
        patients; masks: bool ``[G, T]`` over ``train_use_type`` (True = masked) as a numpy array or a ``MaskPlan``
        (its device-side index tables, which a captured graph can re-read), None = nothing masked.
        Returns a dict of batched tensors: logits_all / logits_<m> ``[G, 4]``, one_x ``[G, 8]``, multi_x
        ``[G, T', 8]``, fea ``[G, T', 512]``, mae_out / mae_labels, att_2 / att_3 (lists of ``[G, nodes]``)."""
        if mix is None:
            mix = self._DEFAULT_MIX
        train_use_type = list(train_use_type)
        use_type = list(train_use_type if use_type is None else use_type)
        Tt = len(train_use_type)
        present = [m for m in MODALITIES if m in use_type]
        G = int(feats[present[0]].shape[0])
        dev = feats[present[0]].device
        C = self.hidden
        P = len(present)
        seg = {m: int(feats[m].shape[1]) for m in present}
        nodes_n = [seg[m] for m in present]
        # every modality's node rows live in ONE stack [sum_m G * nodes_m, C] (modality-major); tab_gm numbers the patient
        # graphs patient-major (pooled rows = the auto-encoder's token order), tab_mg modality-major (pooled rows = the row
        # blocks of the per-modality heads), tab_row is the stack of the P * G pooled rows themselves (segments of one row)
        tab_gm = self._segtab(G, nodes_n, "gm", dev)
        tab_mg = self._segtab(G, nodes_n, "mg", dev)
        tab_row = self._segtab(G, [1] * P, "mg", dev)
        # SAGEConv of every modality (lin_l on the neighbourhood mean, lin_r on the node itself): 2 x P GEMMs, one launch,
        # written straight into the stack
        entries = []
        for i, m in enumerate(present):
            x = feats[m].contiguous().float().reshape(G * seg[m], -1)
            conv = getattr(self, m + "_gnn_2")
            agg = R.GraphMean.apply(x, self._topology(edges[m], seg[m], dev), G)
            r0 = tab_gm.row_start[i]
            entries.append((agg, 0, agg.shape[0], conv.lin_l, "l", r0))
            entries.append((x, 0, x.shape[0], conv.lin_r, "r", r0))
        sage = R.linear_group(entries)
        blocks = [getattr(self, m + "_relu_2") for m in present]
        if len({blk[2].p for blk in blocks}) != 1:
            raise ValueError("the modality branches must share one dropout rate")
        x = ops.relu(ops.Add.apply(sage["l"], sage["r"]))
        x = R.segtab_layernorm(x, tab_gm, [blk[1] for blk in blocks], 0)
        nodes = self._drop(x, blocks[0][2].p)
        pool_x, att = self._pool_stack(nodes, present, "", tab_gm)             # [G * P, C], patient-major
        att_2 = [att[tab_gm.row_start[i]:tab_gm.row_start[i + 1]].reshape(G, seg[m]) for i, m in enumerate(present)]
        out = {"mae_labels": pool_x.reshape(G, len(present), C), "att_2": att_2}
        if Tt > 1:
            plan = None
            if use_type == train_use_type:
                if isinstance(masks, MaskPlan):
                    plan = masks
                    if (plan.G, plan.T) != (G, Tt):
                        raise ValueError("mask plan of shape (%d, %d) for a batch of (%d, %d)" % (plan.G, plan.T, G, Tt))
                else:
                    mk = np.zeros((G, Tt), dtype=bool) if masks is None else np.asarray(masks, dtype=bool).reshape(G, Tt)
                tokens = pool_x
            else:
                # inference with missing modalities (my_mae_model.py:597-612): absent tokens are zero and masked
                slot, k = [], 0
                mk = np.ones((G, Tt), dtype=bool)
                for i, m in enumerate(train_use_type):
                    if m in use_type:
                        slot.append(k); k += 1; mk[:, i] = False
                    else:
                        slot.append(-1)
                if k == 0:
                    mk[:] = False
                idx = [(-1 if s < 0 else g * len(present) + s) for g in range(G) for s in slot]
                tokens = R.RowsGather.apply(pool_x, self._idx(idx, dev), None)
            if plan is None:
                plan = self._plan(mk, dev)
            out["mask_plan"] = plan
            mae_x = self._mae(tokens, plan)
            out["mae_out"] = mae_x.reshape(G, Tt, C)
            out["after_mae"] = mae_x.reshape(G, Tt, C)
            if mix:
                mae_x = self._mixer(mae_x, G, Tt)
                out["after_mix"] = mae_x.reshape(G, Tt, C)
            # node features += the reconstructed token of their modality (my_mae_model.py:636-649), all modalities at once
            tkey = ("tok", G, tuple(present), tuple(train_use_type), str(dev))
            tabs = self._const.get(tkey)
            if tabs is None:
                tok_of_seg = [-1] * (G * P)
                seg_of_tok = [-1] * (G * Tt)
                for i, m in enumerate(present):
                    if m in train_use_type:
                        j = train_use_type.index(m)
                        for g in range(G):
                            tok_of_seg[tab_mg.seg_id(i, g)] = g * Tt + j
                            seg_of_tok[g * Tt + j] = tab_mg.seg_id(i, g)
                tabs = self._const[tkey] = (torch.tensor(tok_of_seg, dtype=torch.int32, device=dev),
                                            torch.tensor(seg_of_tok, dtype=torch.int32, device=dev))
            nodes = R.SegTabBcastAdd.apply(nodes, mae_x, tab_mg, tabs[0], tabs[1])
        pooled2, att = self._pool_stack(nodes, present, "_2", tab_mg)            # [P * G, C], modality-major
        att_3 = [att[tab_mg.row_start[i]:tab_mg.row_start[i + 1]].reshape(G, seg[m]) for i, m in enumerate(present)]
        f = R.L2Norm.apply(pooled2)                                                # F.normalize(dim=1) is row-wise
        # the per-modality heads 512 -> 128 -> 32 -> 8 (-> 4): every layer is one grouped launch over the modalities' row
        # blocks, the LayerNorms (PyG graph mode on a single row) one launch with each block's own affine
        get = lambda name: [getattr(self, name + m) for m in present]           # noqa: E731
        v = ops.relu(self._per_set(f, tab_row, get("lin1_")))
        v = self._drop(R.segtab_layernorm(v, tab_row, get("norm1_"), 0), self.dropout.p)
        v = ops.relu(self._per_set(v, tab_row, get("lin2_")))
        v = self._drop(R.segtab_layernorm(v, tab_row, get("norm2_"), 0), self.dropout.p)
        v = self._per_set(v, tab_row, get("lin3_"))                              # [P * G, 8]
        logits_stack = self._per_set(v, tab_row, get("classifier_"))            # [P * G, 4]
        K = v.shape[1]
        if P > 1:       # modality-major -> patient-major for the [G, P, .] outputs and the mean over the modalities
            to_gm = self._idx([i * G + g for g in range(G) for i in range(P)], dev)
            multi_x = R.RowsGather.apply(v, to_gm, None).reshape(G, P, 1, K)
            one_x = ops.global_avg_pool(multi_x).reshape(G, K)
            fea = R.RowsGather.apply(f, to_gm, None).reshape(G, P, C)
        else:
            multi_x, one_x, fea = v.reshape(G, 1, 1, K), v.reshape(G, K), f.reshape(G, 1, C)
        out.update(one_x=one_x, multi_x=multi_x.reshape(G, P, K), logits_all=R.linear(one_x, self.classifier),
                   att_3=att_3, fea=fea, present=present, logits_stack=logits_stack)
        for i, m in enumerate(present):
            out["logits_" + m] = logits_stack[i * G:(i + 1) * G]
        return out

    # ------------------------------------------------------------------ reference signature (one patient)
    def forward(self, all_thing, train_use_type=None, use_type=None, in_mask=[], mix=None):
        if mix is None:
            mix = self._DEFAULT_MIX
        get = (lambda k: all_thing[k]) if isinstance(all_thing, dict) else (lambda k: getattr(all_thing, k))
        train_use_type = list(train_use_type)
        use_type = list(train_use_type if use_type is None else use_type)
        dev = self.classifier.weight.device
        feats = {m: torch.as_tensor(get("x_" + m)).to(dev).unsqueeze(0) for m in MODALITIES if m in use_type}
        edges = {m: get(_EDGE_ATTR[m]) for m in MODALITIES if m in use_type}
        masks = None if len(in_mask) == 0 else np.asarray(in_mask, dtype=bool).reshape(1, -1)
        o = self.forward_batch(feats, edges, train_use_type, use_type, masks, mix)
        save_fea, fea_dict = {}, {"mae_labels": o["mae_labels"][0]}
        if "mae_out" in o:
            fea_dict["mae_out"] = o["mae_out"][0]
            save_fea["after_mae"] = o["after_mae"][0].detach().cpu().numpy()
            if "after_mix" in o:
                save_fea["after_mix"] = o["after_mix"][0].detach().cpu().numpy()
        for k, m in enumerate(o["present"]):
            fea_dict[m] = o["fea"][0, k]
        lg = {m: (o["logits_" + m][0] if m in o["present"] else None) for m in MODALITIES}
        att_2 = [a.reshape(-1, 1) for a in o["att_2"]]
        att_3 = [a.reshape(-1, 1) for a in o["att_3"]]
        head = ((o["one_x"][0], o["multi_x"][0]), save_fea, (att_2, att_3), fea_dict, o["logits_all"][0])
        if self._RETURN_LOGITS == "img+cli":   # my_mae_model_2.py:771-779: the one image modality present, then cli
            img = None
            for m in ("imgN", "imgA", "imgL"):
                if lg[m] is not None:
                    img = lg[m]                # the reference overwrites img_logits in this order (:230-260)
            return head + (img, lg["cli"])
        return head + tuple(lg[m] for m in self._RETURN_LOGITS)


# ----------------------------------------------------------------------------- the reference's per-variant classes
class fusion_model_mae_three(fusion_model_mae_2):
    """``Three_Modal/my_mae_model_three.py``: 3 training modalities by default, ``mix=False`` by default, the unused
    ``norm3_*`` layers, the four-modal 9-tuple (absent modalities' logits are None)."""
    _DEFAULT_TYPES, _DEFAULT_MIX, _NORM3 = 3, False, True


class fusion_model_mae_two(fusion_model_mae_2):
    """``Two_Modal/my_mae_model_2.py`` (one image modality + clinical: train(NC|AC|LC).py): 7-tuple ending in
    (all_logits, img_logits, cli_logits)."""
    _DEFAULT_TYPES, _DEFAULT_MIX, _NORM3 = 2, False, True
    _RETURN_LOGITS = "img+cli"


class fusion_model_mae_two_NL(fusion_model_mae_two):
    """``Two_Modal/my_mae_model_2_NL.py``: (all_logits, imgN_logits, imgL_logits)."""
    _RETURN_LOGITS = ("imgN", "imgL")


class fusion_model_mae_two_AL(fusion_model_mae_two):
    """``Two_Modal/my_mae_model_2_AL.py``: (all_logits, imgA_logits, imgL_logits)."""
    _RETURN_LOGITS = ("imgA", "imgL")


class fusion_model_mae_two_NA(fusion_model_mae_two):
    """``Two_Modal/my_mae_model_2_NA.py``: (all_logits, imgN_logits, imgA_logits)."""
    _RETURN_LOGITS = ("imgN", "imgA")


# ----------------------------------------------------------------------------- objective + train step
MODALITY_LOSS_WEIGHT = {"imgN": 0.3, "imgA": 0.3, "imgL": 0.3, "cli": 0.2}


def fusion_objective(out: Dict[str, torch.Tensor], labels: torch.Tensor, masks=None, mse_factor: float = 5.0):
    """my_train(full).py:309-347 for a batch of G patients: CE(all) + 0.3 CE(img*) + 0.2 CE(cli) on the stacked
    ``[G, 4]`` logits + sum_g mse_factor * MSE(mae_out[g][masked], pool_x[g][masked]) / G / 5."""
    present = out["present"]
    G, T, C = out["mae_out"].shape
    # the masks are the ones forward_batch was given: their device-side tables travel in ``out`` (no host work here)
    plan = masks if isinstance(masks, MaskPlan) else out.get("mask_plan")
    if plan is None:
        plan = MaskPlan(np.asarray(masks, dtype=bool).reshape(G, T), labels.device)
    n_masked = T - plan.n_vis
    sel = plan.sel
    weights = [1.0] + [MODALITY_LOSS_WEIGHT[m] for m in present]
    inv_count = 1.0 / max(n_masked * C, 1)
    return R.FusionObjective.apply(labels, sel, weights, mse_factor / G / 5.0, inv_count,
                                   out["mae_out"].reshape(G * T, C), out["mae_labels"].reshape(G * T, C),
                                   out["logits_all"], out["logits_stack"])


def generate_mask(num=3):
    """mae_utils.py:11-21: exactly one visible modality; shape [1,1,num] bool (True = masked)."""
    mask = np.hstack([np.zeros(1, dtype=bool), np.ones(num - 1, dtype=bool)])
    np.random.shuffle(mask)
    return mask[None, None, :]


def get_edge_index_image():
    """4x4 patch grid, 8-neighbourhood, both directions (Graph_Structure(data_augmentation).py:338-355) -> [2, 84]."""
    start, end = [], []
    for pos in range(16):
        r, c = divmod(pos, 4)
        for rr in range(max(r - 1, 0), min(r + 2, 4)):
            for cc in range(max(c - 1, 0), min(c + 2, 4)):
                if (rr, cc) != (r, c):
                    start.append(pos); end.append(rr * 4 + cc)
    return torch.tensor([start, end], dtype=torch.long)


def get_edge_index_full(n=4):
    """complete digraph on the clinical nodes (util.py:69-77, Graph_Structure...py:367-376) -> [2, n(n-1)]."""
    start, end = [], []
    for i in range(n):
        for j in range(n):
            if i != j:
                start.append(j); end.append(i)
    return torch.tensor([start, end], dtype=torch.long)
