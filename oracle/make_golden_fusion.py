"""Generate tests/golden/fusion_*.npz by running the reference's UNMODIFIED fusion model
(MultiModal Prediction/Four_Modal/my_mae_model.py) on top of the dependency stand-ins in
oracle/_shims (see its README for what is and is not pinned), then assert that the oracle
restatement (oracle/fusion_ref.py) agrees.  Build container only:  python oracle/make_golden_fusion.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fusion_ref as FR  # noqa: E402

REF_DIR = "/root/reference/MultiModal Prediction/Four_Modal"


class legacy_index:
    """torch < 2.9 treated a numpy bool index of shape [1,1,T] on a [1,T,C] tensor as the tuple
    (mask[0],) i.e. x[:, mask[0,0]] after broadcasting; torch 2.11 raises IndexError
    (my_mae_model.py:143,318-319; my_train(full).py:253).  This adapter restores that reading
    WITHOUT touching the reference source."""

    def __enter__(self):
        self.orig = torch.Tensor.__getitem__

        def getitem(t, idx):
            if isinstance(idx, np.ndarray) and idx.dtype == bool and idx.ndim == 3 and idx.shape[:2] == (1, 1):
                return self.orig(t, (slice(None), torch.from_numpy(idx[0, 0])))
            return self.orig(t, idx)

        torch.Tensor.__getitem__ = getitem

    def __exit__(self, *exc):
        torch.Tensor.__getitem__ = self.orig


def import_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_shims"))
    sys.path.insert(0, REF_DIR)
    import my_mae_model as M
    from torch_geometric.data import Data
    sys.path.remove(REF_DIR)
    return M, Data


def to_data(Data, g, use_types):
    return Data(x_imgN=g["x_imgN"], x_imgA=g["x_imgA"], x_imgL=g["x_imgL"], x_cli=g["x_cli"], data_id="p",
                data_type=list(use_types), edge_index_imageN=g["edge_index_imageN"],
                edge_index_imageA=g["edge_index_imageA"], edge_index_imageL=g["edge_index_imageL"],
                edge_index_cli=g["edge_index_cli"])


def main():
    M, Data = import_reference()
    M.device = torch.device("cpu")
    out_dir = os.path.join(ROOT, "tests", "golden")
    for use_types, tag in ((["imgN", "imgA", "imgL", "cli"], "4modal"), (["imgN", "imgA", "imgL"], "3modal")):
        T = len(use_types)
        torch.manual_seed(0)
        ref = M.fusion_model_mae_2(1024, 512, 512, 0.3, T)
        state = FR.randomize_state(ref.state_dict(), seed=3)
        ref.load_state_dict(state, strict=True)
        ref.eval()                                   # dropout off: parity of values and gradients
        masks = [np.array([[[True] * (T - 1) + [False]]]), np.array([[[False, True] + [True] * (T - 2)]])]
        labels = torch.tensor([2, 0])
        payload = {"use_types": np.array(use_types), "labels": labels.numpy(),
                   "masks": np.stack([m[0, 0] for m in masks])}
        logits = {k: [] for k in ["all"] + use_types}
        mse = 0.0
        outs_or = []
        st = {k: v.clone().requires_grad_(True) for k, v in state.items()}
        for i in range(2):
            g = FR.synthetic_patient(i)
            with legacy_index():
                res = ref(to_data(Data, g, use_types), use_types, use_types, masks[i], mix=True)
                (one_x, multi_x), _, (att2, att3), fea, l_all, l_N, l_A, l_L, l_cli = res
                mk = masks[i]
                mse = mse + 5.0 * torch.nn.functional.mse_loss(fea["mae_out"][mk[0]] if False else fea["mae_out"][torch.from_numpy(mk[0, 0])],
                                                               fea["mae_labels"][torch.from_numpy(mk[0, 0])])
            per = {"imgN": l_N, "imgA": l_A, "imgL": l_L, "cli": l_cli}
            logits["all"].append(l_all)
            for m in use_types:
                logits[m].append(per[m])
            o = FR.fusion_forward(g, st, use_types, masks[i][0, 0], mix=True)
            outs_or.append(o)
            for name, a, b in (("one_x", one_x, o["one_x"]), ("multi_x", multi_x, o["multi_x"]), ("logits_all", l_all, o["logits_all"]),
                               ("mae_out", fea["mae_out"], o["mae_out"]), ("mae_labels", fea["mae_labels"], o["mae_labels"]),
                               ("att3_0", att3[0], o["att_3"][0])):
                err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
                assert err < 2e-5, (tag, i, name, err)
                payload["p%d:%s" % (i, name)] = a.detach().numpy()
            for m in use_types:
                payload["p%d:logits_%s" % (i, m)] = per[m].detach().numpy()
        w = {"imgN": 0.3, "imgA": 0.3, "imgL": 0.3, "cli": 0.2}
        ce = torch.nn.CrossEntropyLoss()
        loss = ce(torch.stack(logits["all"]), labels)
        for m in use_types:
            loss = loss + w[m] * ce(torch.stack(logits[m]), labels)
        loss = loss + mse / 2 / 5
        loss.backward()
        loss_or = FR.fusion_loss(outs_or, [m[0, 0] for m in masks], labels, use_types)
        loss_or.backward()
        assert abs(float(loss) - float(loss_or)) < 1e-5 * abs(float(loss)), (float(loss), float(loss_or))
        worst = 0.0
        ref_grads = dict(ref.named_parameters())
        for k, p in ref_grads.items():
            if p.grad is None:
                assert st[k].grad is None or float(st[k].grad.abs().max()) == 0, k
                continue
            if k.endswith("gate_nn.2.bias"):   # softmax over the nodes is shift invariant: exactly-zero gradient
                assert float(p.grad.abs().max()) < 1e-6 and float(st[k].grad.abs().max()) < 1e-6, k
                continue
            e = float((st[k].grad - p.grad).abs().max() / p.grad.abs().max().clamp_min(1e-6))  # softmax-shift biases have ~0 grads
            if e > 5e-4:
                print("   ", k, e, float(p.grad.abs().max()))
            worst = max(worst, e)
        print("%s: loss %.6f (oracle %.6f), worst grad rel err %.2e over %d tensors" % (tag, float(loss), float(loss_or), worst, len(ref_grads)))
        assert worst < 5e-4, worst
        payload["loss"] = float(loss)
        for k in ("imgN_gnn_2.lin_l.weight", "mpool_imgA.gate_nn.0.weight", "mae.encoder.blocks.0.attn.qkv.weight",
                  "mae.decoder.head.bias", "mae.mask_token", "mix.mix_mip_1.0.weight", "mix.norm.weight", "lin2_imgL.weight",
                  "classifier.weight", "imgL_relu_2.1.bias"):
            gk = ref_grads[k].grad
            payload["grad:" + k] = gk.reshape(-1)[:: gk.numel() // 8192 + 1].numpy()
        payload["state_keys"] = np.array(list(state.keys()))
        np.savez_compressed(os.path.join(out_dir, "fusion_%s.npz" % tag), **payload)
    variants(out_dir, Data)
    print("fusion golden vectors written")


def load_variant(subdir, fname):
    """Import one of the reference's per-variant copies (Two_Modal/my_mae_model_2*.py, Three_Modal/my_mae_model_three.py)
    unmodified, with ITS directory's mae_utils / util on the path."""
    import importlib.util
    d = os.path.join("/root/reference/MultiModal Prediction", subdir)
    for k in ("mae_utils", "util"):
        sys.modules.pop(k, None)
    sys.path.insert(0, d)
    try:
        spec = importlib.util.spec_from_file_location("ref_" + fname[:-3], os.path.join(d, fname))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(d)
    mod.device = torch.device("cpu")
    return mod


VARIANTS = (  # tag, directory, file, modalities of the train script, names of the logits the 7-/9-tuple ends with
    ("two_NC", "Two_Modal", "my_mae_model_2.py", ["imgN", "cli"], ["img", "cli"]),
    ("two_LC", "Two_Modal", "my_mae_model_2.py", ["imgL", "cli"], ["img", "cli"]),
    ("two_NL", "Two_Modal", "my_mae_model_2_NL.py", ["imgN", "imgL"], ["imgN", "imgL"]),
    ("two_AL", "Two_Modal", "my_mae_model_2_AL.py", ["imgA", "imgL"], ["imgA", "imgL"]),
    ("two_NA", "Two_Modal", "my_mae_model_2_NA.py", ["imgN", "imgA"], ["imgN", "imgA"]),
    ("three_NAL", "Three_Modal", "my_mae_model_three.py", ["imgN", "imgA", "imgL"], ["imgN", "imgA", "imgL", "cli"]),
    ("three_NLC", "Three_Modal", "my_mae_model_three.py", ["imgN", "imgL", "cli"], ["imgN", "imgA", "imgL", "cli"]),
)


def variants(out_dir, Data):
    """tests/golden/fusion_variants.npz: for every per-variant model file of the reference its state_dict key list and
    one patient's forward with the file's OWN defaults (train_type_num, mix), checked against the oracle."""
    payload = {}
    for tag, subdir, fname, use_types, tail in VARIANTS:
        M = load_variant(subdir, fname)
        T = len(use_types)
        torch.manual_seed(0)
        ref = M.fusion_model_mae_2(1024, 512, 512, 0.3)          # the file's default train_type_num
        assert ref.train_type_num == T if hasattr(ref, "train_type_num") else True
        state = FR.randomize_state(ref.state_dict(), seed=5)
        ref.load_state_dict(state, strict=True)
        ref.eval()
        mask = np.array([[[False] + [True] * (T - 1)]])
        g = FR.synthetic_patient(7)
        with legacy_index(), torch.no_grad():
            res = ref(to_data(Data, g, use_types), use_types, use_types, mask)      # mix: the file's default (False)
        (one_x, multi_x), _, (att2, att3), fea, l_all = res[:5]
        assert len(res) == 5 + len(tail), (tag, len(res))
        o = FR.fusion_forward(g, {k: v for k, v in state.items() if not k.startswith("norm3_")}, use_types, mask[0, 0],
                              mix=False)
        for name, a, b in (("one_x", one_x, o["one_x"]), ("multi_x", multi_x, o["multi_x"]), ("logits_all", l_all, o["logits_all"]),
                           ("mae_out", fea["mae_out"], o["mae_out"])):
            err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
            assert err < 2e-5, (tag, name, err)
            payload["%s:%s" % (tag, name)] = a.detach().numpy()
        for i, name in enumerate(tail):
            v = res[5 + i]
            payload["%s:tail%d" % (tag, i)] = np.zeros(0, dtype=np.float32) if v is None else v.detach().numpy()
        payload["%s:state_keys" % tag] = np.array(list(state.keys()))
        payload["%s:use_types" % tag] = np.array(use_types)
        payload["%s:mask" % tag] = mask[0, 0]
        print("%s (%s/%s): %d state entries, %d-tuple, oracle agrees" % (tag, subdir, fname, len(state), len(res)))
    payload["tags"] = np.array([v[0] for v in VARIANTS])
    payload["files"] = np.array([v[2] for v in VARIANTS])
    np.savez_compressed(os.path.join(out_dir, "fusion_variants.npz"), **payload)


if __name__ == "__main__":
    main()
