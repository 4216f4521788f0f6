"""Shared body of the fusion-head parity tests (CPU: host graph over the emulated C ABI; GPU: the CUDA kernels).
Golden vectors: tests/golden/fusion_{4,3}modal.npz, produced by oracle/make_golden_fusion.py from the reference's
unmodified my_mae_model.py (dependency stand-ins: oracle/_shims)."""
import os

import numpy as np
import torch

from cervix_b200.multimodal.my_mae_model import fusion_model_mae_2, fusion_objective
from oracle import fusion_ref as FR

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relerr(a, b):
    a = torch.as_tensor(a).detach().float().cpu(); b = torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def build(tag, device):
    g = np.load(os.path.join(GOLDEN, "fusion_%s.npz" % tag))
    use_types = [str(u) for u in g["use_types"]]
    model = fusion_model_mae_2(1024, 512, 512, 0.3, len(use_types))
    assert list(model.state_dict().keys()) == [str(k) for k in g["state_keys"]]
    model.load_state_dict(FR.randomize_state(model.state_dict(), seed=3), strict=True)
    model.to(device).eval()       # golden vectors were produced with dropout off
    return g, use_types, model


def batch_of(patients, use_types, device):
    feats = {m: torch.stack([p["x_" + m] for p in patients]).to(device) for m in use_types}
    edge_key = {"imgN": "edge_index_imageN", "imgA": "edge_index_imageA", "imgL": "edge_index_imageL", "cli": "edge_index_cli"}
    edges = {m: patients[0][edge_key[m]] for m in use_types}
    return feats, edges


def check_batched(tag, device, tol=2e-4, gtol=2e-3):
    g, use_types, model = build(tag, device)
    patients = [FR.synthetic_patient(i) for i in range(2)]
    feats, edges = batch_of(patients, use_types, device)
    masks = g["masks"]
    out = model.forward_batch(feats, edges, use_types, use_types, masks, True)
    for i in range(2):
        for name in ("one_x", "multi_x", "logits_all", "mae_out", "mae_labels"):
            assert relerr(out[name][i], g["p%d:%s" % (i, name)]) < tol, (tag, i, name)
        assert relerr(out["att_3"][0][i].reshape(-1, 1), g["p%d:att3_0" % i]) < tol
        for m in use_types:
            assert relerr(out["logits_" + m][i], g["p%d:logits_%s" % (i, m)]) < tol, (tag, i, m)
    labels = torch.from_numpy(g["labels"]).to(device)
    loss = fusion_objective(out, labels, masks)
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"])), (float(loss), float(g["loss"]))
    loss.backward()
    params = dict(model.named_parameters())
    worst = 0.0
    for key in g.files:
        if not key.startswith("grad:"):
            continue
        gk = params[key[5:]].grad
        assert gk is not None, key
        got = gk.reshape(-1)[:: gk.numel() // 8192 + 1]
        e = relerr(got, g[key])
        worst = max(worst, e)
        assert e < gtol, (tag, key, e)
    # parameters the reference never reaches stay gradient-free
    assert params["fc_cli_1.weight"].grad is None
    return worst


def check_single(tag, device, tol=2e-4):
    """The reference's one-patient signature and 9-tuple (my_mae_model.py:500, 783-793)."""
    g, use_types, model = build(tag, device)
    p = FR.synthetic_patient(0)
    p = dict(p, data_id="p", data_type=use_types)
    with torch.no_grad():
        res = model(p, use_types, use_types, g["masks"][0][None, None, :], mix=True)
    (one_x, multi_x), save_fea, (att_2, att_3), fea, l_all, l_n, l_a, l_l, l_c = res
    assert relerr(one_x, g["p0:one_x"]) < tol and relerr(multi_x, g["p0:multi_x"]) < tol
    assert relerr(l_all, g["p0:logits_all"]) < tol
    per = {"imgN": l_n, "imgA": l_a, "imgL": l_l, "cli": l_c}
    for m in ("imgN", "imgA", "imgL", "cli"):
        if m in use_types:
            assert relerr(per[m], g["p0:logits_" + m]) < tol
        else:
            assert per[m] is None
    assert relerr(fea["mae_out"], g["p0:mae_out"]) < tol and relerr(fea["mae_labels"], g["p0:mae_labels"]) < tol
    assert isinstance(save_fea["after_mae"], np.ndarray) and save_fea["after_mix"].shape == (len(use_types), 512)
    assert att_3[0].shape == (16, 1) and relerr(att_3[0], g["p0:att3_0"]) < tol
    assert len(att_2) == len(use_types) and fea["imgN"].shape == (512,)


def check_missing_modality(device):
    """Inference with a modality absent (use_type != train_use_type, my_mae_model.py:597-612) against the oracle
    restatement of the same branch: absent tokens enter the MAE as masked zeros."""
    g, use_types, model = build("4modal", device)
    present = ["imgN", "imgL", "cli"]
    p = dict(FR.synthetic_patient(1), data_id="p", data_type=present)
    with torch.no_grad():
        res = model(p, use_types, present, [], mix=True)
    (one_x, multi_x), _, _, fea, l_all, l_n, l_a, l_l, l_c = res
    assert l_a is None and multi_x.shape == (3, 8) and fea["mae_out"].shape == (4, 512)
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        ref = oracle_missing(p, state, use_types, present)
    assert relerr(fea["mae_out"], ref["mae_out"]) < 2e-4
    assert relerr(l_all, ref["logits_all"]) < 2e-4


def oracle_missing(graph, state, train_types, present):
    import torch.nn.functional as F
    edges = {"imgN": graph["edge_index_imageN"], "imgA": graph["edge_index_imageA"],
             "imgL": graph["edge_index_imageL"], "cli": graph["edge_index_cli"]}
    nodes, pooled = {}, {}
    for m in present:
        x = FR.gnn_relu_block(FR.sage_conv(graph["x_" + m], edges[m], state, m + "_gnn_2"), state, m + "_relu_2")
        nodes[m] = x
        pooled[m] = FR.gate_pool(x, state, "mpool_" + m)[0]
    tokens = torch.cat([pooled[m] if m in present else torch.zeros(1, 512) for m in train_types], 0)
    mask = [m not in present for m in train_types]
    mae_out = FR.mae_forward(tokens, mask, state)
    mae_x = FR.mixer_block(mae_out, state)
    feats = []
    for m in present:
        nodes[m] = nodes[m] + mae_x[train_types.index(m)]
        feats.append(FR.gate_pool(nodes[m], state, "mpool_" + m + "_2")[0])
    x = F.normalize(torch.cat(feats, 0), dim=1)
    multi = torch.stack([FR._head(x[i], state, m)[0] for i, m in enumerate(present)])
    return {"mae_out": mae_out, "logits_all": FR._lin(multi.mean(0), state, "classifier")}


# ---- the reference's per-variant model files (Two_Modal/my_mae_model_2*.py, Three_Modal/my_mae_model_three.py) -------
VARIANT_CLASS = {"my_mae_model_2.py": "fusion_model_mae_two", "my_mae_model_2_NL.py": "fusion_model_mae_two_NL",
                 "my_mae_model_2_AL.py": "fusion_model_mae_two_AL", "my_mae_model_2_NA.py": "fusion_model_mae_two_NA",
                 "my_mae_model_three.py": "fusion_model_mae_three"}


def variant_tags():
    g = np.load(os.path.join(GOLDEN, "fusion_variants.npz"))
    return [str(t) for t in g["tags"]]


def check_variant(tag, device, tol=2e-4):
    """One patient through the class that mirrors the variant file, constructed and called with the file's OWN defaults
    (train_type_num, mix), against tests/golden/fusion_variants.npz: same state_dict schema (156 entries with the
    never-used norm3_* layers), same tuple length and logits order, same values."""
    import cervix_b200.multimodal.my_mae_model as MM
    g = np.load(os.path.join(GOLDEN, "fusion_variants.npz"))
    fname = str(g["files"][[str(t) for t in g["tags"]].index(tag)])
    cls = getattr(MM, VARIANT_CLASS[fname])
    use_types = [str(u) for u in g[tag + ":use_types"]]
    model = cls(1024, 512, 512, 0.3)
    assert model.train_type_num == len(use_types)
    keys = [str(k) for k in g[tag + ":state_keys"]]
    assert list(model.state_dict().keys()) == keys and len(keys) == 156
    model.load_state_dict(FR.randomize_state(model.state_dict(), seed=5), strict=True)
    model.to(device).eval()
    p = dict(FR.synthetic_patient(7), data_id="p", data_type=use_types)
    with torch.no_grad():
        res = model(p, use_types, use_types, g[tag + ":mask"][None, None, :])
    n_tail = len([k for k in g.files if k.startswith(tag + ":tail")])
    assert len(res) == 5 + n_tail
    (one_x, multi_x), save_fea, _, fea, l_all = res[:5]
    assert "after_mix" not in save_fea                       # mix defaults to False in these files
    for name, got in (("one_x", one_x), ("multi_x", multi_x), ("logits_all", l_all), ("mae_out", fea["mae_out"])):
        assert relerr(got, g["%s:%s" % (tag, name)]) < tol, (tag, name)
    for i in range(n_tail):
        want = g["%s:tail%d" % (tag, i)]
        if want.size == 0:
            assert res[5 + i] is None, (tag, i)
        else:
            assert relerr(res[5 + i], want) < tol, (tag, i)
