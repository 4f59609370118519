"""Helpers to load the committed golden vectors (minted by tests/golden/make_golden.py from the
unmodified reference) as torch tensors."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("case_eval_alias_k2", "case_train_rand_k3", "case_train_empty_k2")
GRAPH_KEYS = ("ui", "iu", "img_ui", "img_iu", "txt_ui", "txt_iu")


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.z = z
        self.name = name
        self.cfg = json.loads(str(z["cfg"]))
        self.params = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p/")}
        self.grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
        self.losses = {k[5:]: float(z[k]) for k in z.files if k.startswith("loss/")}
        self.outs = [torch.from_numpy(z[f"out{j}"]) for j in range(12)]
        self.image_feats = torch.from_numpy(z["image_feats"])
        self.text_feats = torch.from_numpy(z["text_feats"])
        self.masks = (torch.from_numpy(z["mask0"]), torch.from_numpy(z["mask1"]))
        self.users, self.pos, self.neg = (torch.from_numpy(z[k]).long() for k in ("users", "pos", "neg"))

    def graphs(self, device="cpu"):
        """The six torch sparse COO tensors; aliasing between modality and ui/iu graphs is
        restored for the 'alias' case (Trainer.__init__, main.py:68-69)."""
        out = {}
        for k in GRAPH_KEYS:
            idx = torch.from_numpy(self.z[f"g_{k}_idx"])
            val = torch.from_numpy(self.z[f"g_{k}_val"])
            shape = tuple(int(s) for s in self.z[f"g_{k}_shape"])
            out[k] = torch.sparse_coo_tensor(idx, val, shape).to(device)
        if self.cfg["modal"] == "alias":
            out["img_ui"] = out["txt_ui"] = out["ui"]
            out["img_iu"] = out["txt_iu"] = out["iu"]
        elif self.cfg["modal"] == "empty":
            out["txt_ui"], out["txt_iu"] = out["img_ui"], out["img_iu"]
        return [out[k] for k in GRAPH_KEYS]

    def oracle_cfg(self):
        from oracle.mmssl_oracle import HotPathConfig
        c = self.cfg
        return HotPathConfig(embed_size=c["d"], n_layers=c["n_layers"], head_num=c["head_num"],
                             id_cat_rate=c["id_cat_rate"], model_cat_rate=c["model_cat_rate"],
                             drop_rate=c["drop_rate"], tau=c["tau"], cl_rate=c["cl_rate"],
                             emb_decay=c["emb_decay"], feat_reg_decay=c["feat_reg_decay"], batch_size=c["B"])

    @property
    def train(self):
        return bool(self.cfg["train"])


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max(|b|_inf, tiny): the 'relative fp32' measure used for the 1e-4 contract."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = max(float(b.abs().max()), 1e-30)
    return float((a - b).abs().max()) / denom


def elem_err(a: torch.Tensor, b: torch.Tensor, rtol: float = 1e-4, atol_frac: float = 1e-6) -> float:
    """Per-element form of the 1e-4 contract: max over elements of |a-b| / (rtol*|b| + atol_frac*max|b|); <= 1 passes.
    (rel_err above is norm-wise: one large entry hides the small ones.  The absolute floor is fp32 noise of the reference's
    own sums: 1e-6 of the tensor's largest magnitude.)"""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    floor = atol_frac * max(float(b.abs().max()), 1e-30)
    return float(((a - b).abs() / (rtol * b.abs() + floor)).max()) if b.numel() else 0.0
