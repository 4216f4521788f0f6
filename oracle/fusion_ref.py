"""CPU oracle for the multimodal fusion head (TEST INFRASTRUCTURE ONLY; import rules as in
oracle/deeplab_ref.py).

Functional restatement, in plain torch fp32, of ``fusion_model_mae_2.forward``
(/root/reference/MultiModal Prediction/Four_Modal/my_mae_model.py:500-793) and the pieces it calls:
  * SAGEConv (PyG)                         -> sage_conv          (my_mae_model.py:404-416, 544)
  * GNN_relu_Block = ReLU, PyG LayerNorm, Dropout -> gnn_relu_block (:385-397)
  * my_GlobalAttention                     -> gate_pool          (:35-63)
  * PretrainVisionTransformer (MAE)        -> mae_forward        (:216-335, mae_utils.py:58-134)
  * MixerBlock                             -> mixer_block        (:345-369)
  * per-modality heads + classifiers       -> fusion_forward     (:706-793)
  * training objective                     -> fusion_loss        (my_train(full).py:233-253,309-347)

PARITY STATUS: the reference needs torch_geometric / torch_scatter / timm, which are neither vendored
nor version-pinned nor installed here.  ``oracle/make_golden_fusion.py`` runs the reference's UNMODIFIED
``my_mae_model.py`` on top of minimal stand-ins (oracle/_shims) that restate the published algorithm of
each library call, and this file is validated against those outputs (tests/golden/fusion_*.npz).  The
semantics of the three libraries themselves are therefore "parity unpinned" (SURVEY.md section 8c).

All functions take a flat ``state`` dict keyed like the reference's ``state_dict()`` (148 entries).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]
MODALITIES = ("imgN", "imgA", "imgL", "cli")

# 4x4 patch grid, 8-neighbourhood (Graph_Structure(data_augmentation).py:338-355)
GRID_ADJ = {
    0: [1, 4, 5], 1: [0, 2, 4, 5, 6], 2: [1, 3, 5, 6, 7], 3: [2, 6, 7],
    4: [0, 1, 5, 8, 9], 5: [0, 1, 2, 4, 6, 8, 9, 10], 6: [1, 2, 3, 5, 7, 9, 10, 11], 7: [2, 3, 6, 10, 11],
    8: [4, 5, 9, 12, 13], 9: [4, 5, 6, 8, 10, 12, 13, 14], 10: [5, 6, 7, 9, 11, 13, 14, 15], 11: [6, 7, 10, 14, 15],
    12: [8, 9, 13], 13: [8, 9, 10, 12, 14], 14: [9, 10, 11, 13, 15], 15: [10, 11, 14],
}


def image_edge_index() -> torch.Tensor:
    start, end = [], []
    for pos in range(16):
        for nb in GRID_ADJ[pos]:
            start.append(pos); end.append(nb)
    return torch.tensor([start, end], dtype=torch.long)          # [2, 84]


def cli_edge_index(n: int = 4) -> torch.Tensor:
    start, end = [], []
    for i in range(n):
        for j in range(n):
            if i != j:
                start.append(j); end.append(i)
    return torch.tensor([start, end], dtype=torch.long)          # [2, 12]


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    pos = np.arange(n_position)[:, None].astype(np.float64)
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.tensor(table, dtype=torch.float32)              # [T, d]


# ----------------------------------------------------------------------------- building blocks
def _lin(x, state, p, bias=True):
    return F.linear(x, state[p + ".weight"], state[p + ".bias"] if bias and (p + ".bias") in state else None)


def sage_conv(x, edge_index, state, p):
    src, dst = edge_index[0], edge_index[1]
    agg = torch.zeros_like(x).index_add(0, dst, x[src])
    deg = torch.zeros(x.shape[0], dtype=x.dtype).index_add(0, dst, torch.ones(dst.shape[0], dtype=x.dtype))
    agg = agg / deg.clamp_min(1).unsqueeze(-1)
    return _lin(agg, state, p + ".lin_l") + F.linear(x, state[p + ".lin_r.weight"])


def graph_layernorm(x, w, b, eps=1e-5):
    """PyG LayerNorm(mode='graph'): statistics over ALL elements of x, eps added to the std."""
    x = x - x.mean()
    return x / (x.std(unbiased=False) + eps) * w + b


def gnn_relu_block(x, state, p, drop=0.0, training=False):
    x = graph_layernorm(F.relu(x), state[p + ".1.weight"], state[p + ".1.bias"])
    return F.dropout(x, drop, training)


def gate_pool(x, state, p):
    """my_GlobalAttention with a single graph: softmax over the nodes of gate_nn(x), weighted sum."""
    g = _lin(F.relu(_lin(x, state, p + ".gate_nn.0")), state, p + ".gate_nn.2")       # [n,1]
    g = g - g.max()
    e = g.exp()
    att = e / (e.sum() + 1e-16)
    return (att * x).sum(0, keepdim=True), att


def _attention(x, state, p, heads):
    B, N, C = x.shape
    qkv = F.linear(x, state[p + ".qkv.weight"])
    hd = qkv.shape[-1] // (3 * heads)
    qkv = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)).softmax(-1)
    y = (attn @ v).transpose(1, 2).reshape(B, N, heads * hd)
    return _lin(y, state, p + ".proj")


def _block(x, state, p, heads):
    x = x + _attention(F.layer_norm(x, (x.shape[-1],), state[p + ".norm1.weight"], state[p + ".norm1.bias"]),
                       state, p + ".attn", heads)
    h = F.layer_norm(x, (x.shape[-1],), state[p + ".norm2.weight"], state[p + ".norm2.bias"])
    h = _lin(F.gelu(_lin(h, state, p + ".mlp.fc1")), state, p + ".mlp.fc2")
    return x + h


def mae_forward(tokens, mask: Sequence[bool], state, p="mae"):
    """tokens [T,C]; mask[t] True = masked.  Returns [T,C] in modality order (eval-mode: no dropout)."""
    T, C = tokens.shape
    pos = sinusoid_table(T, C)
    mask_t = torch.as_tensor(np.asarray(mask, dtype=bool))
    x = _lin(tokens, state, p + ".encoder.patch_embed") + pos
    vis = x[~mask_t].unsqueeze(0)
    vis = _block(vis, state, p + ".encoder.blocks.0", 12)
    vis = F.layer_norm(vis, (C,), state[p + ".encoder.norm.weight"], state[p + ".encoder.norm.bias"])
    vis = F.linear(vis, state[p + ".encoder_to_decoder.weight"])
    full = torch.cat([vis + pos[~mask_t].unsqueeze(0),
                      state[p + ".mask_token"] + pos[mask_t].unsqueeze(0)], dim=1)
    full = _block(full, state, p + ".decoder.blocks.0", 8)
    full = _lin(F.layer_norm(full, (C,), state[p + ".decoder.norm.weight"], state[p + ".decoder.norm.bias"]),
                state, p + ".decoder.head")[0]
    out = torch.zeros_like(full)
    n_vis = int((~mask_t).sum())
    vi, mi = 0, 0
    for t in range(T):
        if mask_t[t]:
            out[t] = full[n_vis + mi]; mi += 1
        else:
            out[t] = full[vi]; vi += 1
    return out


def mixer_block(x, state, p="mix"):
    w, b = state[p + ".norm.weight"], state[p + ".norm.bias"]
    y = graph_layernorm(x, w, b).t()
    y = _lin(F.gelu(_lin(y, state, p + ".mix_mip_1.0")), state, p + ".mix_mip_1.2").t()
    x = x + y
    y = graph_layernorm(x, w, b)
    return x + _lin(F.gelu(_lin(y, state, p + ".mix_mip_2.0")), state, p + ".mix_mip_2.2")


def _head(v, state, m):
    v = graph_layernorm(F.relu(_lin(v, state, "lin1_" + m)), state["norm1_" + m + ".weight"], state["norm1_" + m + ".bias"])
    v = graph_layernorm(F.relu(_lin(v, state, "lin2_" + m)), state["norm2_" + m + ".weight"], state["norm2_" + m + ".bias"])
    v = _lin(v, state, "lin3_" + m)
    return v, _lin(v, state, "classifier_" + m)


def fusion_forward(graph: dict, state: State, use_types: Sequence[str] = MODALITIES, mask: Sequence[bool] = None,
                   mix: bool = True):
    """One patient, eval-mode semantics (dropout inactive), train_use_type == use_type == ``use_types``.
    graph: {'x_imgN','x_imgA','x_imgL' [16,1024], 'x_cli' [4,1024], 'edge_index_image*' , 'edge_index_cli'}.
    Returns dict(one_x, multi_x, logits_all, logits_<m>, mae_out, mae_labels, att_2, att_3)."""
    T = len(use_types)
    if mask is None:
        mask = [False] * T
    edges = {"imgN": graph["edge_index_imageN"], "imgA": graph["edge_index_imageA"],
             "imgL": graph["edge_index_imageL"], "cli": graph["edge_index_cli"]}
    nodes, pooled, att_2 = {}, [], []
    for m in use_types:
        x = sage_conv(graph["x_" + m], edges[m], state, m + "_gnn_2")
        x = gnn_relu_block(x, state, m + "_relu_2")
        nodes[m] = x
        px, att = gate_pool(x, state, "mpool_" + m)
        pooled.append(px); att_2.append(att)
    pool_x = torch.cat(pooled, 0)                                            # [T,512]
    out = {"mae_labels": pool_x, "att_2": att_2}
    if T > 1:
        mae_x = mae_forward(pool_x, mask, state)
        out["mae_out"] = mae_x
        if mix:
            mae_x = mixer_block(mae_x, state)
        for i, m in enumerate(use_types):
            nodes[m] = nodes[m] + mae_x[i]
    pooled, att_3 = [], []
    for m in use_types:
        px, att = gate_pool(nodes[m], state, "mpool_" + m + "_2")
        pooled.append(px); att_3.append(att)
    x = F.normalize(torch.cat(pooled, 0), dim=1)
    multi = []
    for i, m in enumerate(use_types):
        v, logits = _head(x[i], state, m)
        multi.append(v.unsqueeze(0)); out["logits_" + m] = logits
    multi_x = torch.cat(multi, 0)
    one_x = multi_x.mean(0)
    out.update(one_x=one_x, multi_x=multi_x, logits_all=_lin(one_x, state, "classifier"), att_3=att_3, fea=x)
    return out


def fusion_loss(outs: List[dict], masks: List[Sequence[bool]], labels: torch.Tensor, use_types=MODALITIES,
                mse_factor: float = 5.0):
    """Objective of one reference mini-batch of patients (my_train(full).py:233-253, 309-347)."""
    w = {"imgN": 0.3, "imgA": 0.3, "imgL": 0.3, "cli": 0.2}
    loss = F.cross_entropy(torch.stack([o["logits_all"] for o in outs]), labels)
    for m in use_types:
        loss = loss + w[m] * F.cross_entropy(torch.stack([o["logits_" + m] for o in outs]), labels)
    mse = 0.0
    for o, mk in zip(outs, masks):
        mk = torch.as_tensor(np.asarray(mk, dtype=bool))
        mse = mse + mse_factor * F.mse_loss(o["mae_out"][mk], o["mae_labels"][mk])
    return loss + mse / len(outs) / 5


# ----------------------------------------------------------------------------- synthetic data / weights
def synthetic_patient(seed: int) -> dict:
    g = torch.Generator().manual_seed(4000 + seed)
    return {"x_imgN": torch.randn(16, 1024, generator=g), "x_imgA": torch.randn(16, 1024, generator=g),
            "x_imgL": torch.randn(16, 1024, generator=g), "x_cli": torch.randn(4, 1024, generator=g),
            "edge_index_imageN": image_edge_index(), "edge_index_imageA": image_edge_index(),
            "edge_index_imageL": image_edge_index(), "edge_index_cli": cli_edge_index()}


def randomize_state(state: State, seed: int = 0) -> State:
    """Deterministic non-trivial weights for every entry of a reference-shaped state dict (keeps shapes/keys):
    matrices ~ N(0, 1/fan_in), biases ~ N(0, 0.05), norm weights ~ 1 + N(0, 0.1)."""
    out = {}
    for idx, (k, v) in enumerate(state.items()):
        g = torch.Generator().manual_seed(seed * 7919 + idx)
        if v.dim() >= 2 and "mask_token" not in k:
            t = torch.randn(v.shape, generator=g) / math.sqrt(v.shape[-1])
        elif "mask_token" in k:
            t = 0.02 * torch.randn(v.shape, generator=g)
        elif "norm" in k and k.endswith("weight") or k.endswith(".1.weight"):
            t = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:
            t = 0.05 * torch.randn(v.shape, generator=g)
        out[k] = t
    return out
