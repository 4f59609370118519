"""The UNMODIFIED reference caller driving the drop-in: /root/reference/MMSSL/main.py is imported as
it is, `from Models import MMSSL, Discriminator` (main.py:27) resolves to mmssl_b200/Models.py loaded under the name `Models`
inside the reference tree (so that its `from utility.parser import parse_args` is the reference's own, Models.py:13-15), and
`Trainer.train()` (main.py:308-498) runs three iterations on the cuemu device: the reference's own loop, sampler hook, losses,
Discriminator, both torch optimisers -- with OUR model forward / backward underneath -- and must reproduce what the same loop
recorded with the reference's model (tests/golden/gan_trace.npz: every random draw injected, gradients before and parameters
after each optimiser step of all three iterations).

Build-container only: skipped where /root/reference is absent (the GPU box).  Entry points exercised: main.py:27 (import),
:70-72 (construction, .cuda()), :339-342 (no-grad forward), :363-365 (forward with grad), :368-371, :408-420 (losses on our
outputs, indexed with Python lists), :427-429 (backward through MMSSLForwardFn, AdamW on our parameters)."""
import importlib
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from tests.cuemu import harness
from tests.golden_util import rel_err

REF = "/root/reference/MMSSL"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "main.py")), reason="needs the reference tree (build container only)")

LIVE = ["image_trans.weight", "image_trans.bias", "text_trans.weight", "text_trans.bias", "user_id_embedding.weight",
        "item_id_embedding.weight", "weight_dict.w_self_attention_cat"]


class _Replay(nn.Module):
    """nn.Dropout stand-in that multiplies by the masks the reference run drew (in the order it drew them)."""

    def __init__(self, masks):
        super().__init__()
        self.masks, self.i = masks, 0

    def forward(self, x):
        if not self.training:
            return x
        m = torch.from_numpy(self.masks[self.i])
        self.i += 1
        return x * m


@pytest.fixture
def forget_reference_modules():
    """Modules the test imports from the reference tree (main, utility.*, Models) must not outlive it: a later test's
    `from utility.parser import parse_args` would otherwise get the reference's parser (and its parse of pytest's argv)."""
    before = set(sys.modules)
    yield
    for k in set(sys.modules) - before:
        if k in ("main", "Models") or k.split(".")[0] == "utility":
            del sys.modules[k]


def test_reference_trainer_runs_on_the_drop_in_and_reproduces_its_own_trace(forget_reference_modules, monkeypatch):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import make_dataset
    from make_golden_gan import CASE as c
    z = np.load(os.path.join(ROOT, "tests", "golden", "gan_trace.npz"))
    cfg = json.loads(str(z["cfg"]))
    harness.set_order("fwd")
    harness.emulated_device(monkeypatch)
    tmp = tempfile.mkdtemp(prefix="mmssl_dropin_")
    make_dataset(tmp, c["name"], c["U"], c["I"], c["dv"], c["dt"], c["seed"])
    monkeypatch.syspath_prepend(REF)
    monkeypatch.chdir(REF)
    monkeypatch.setattr(sys, "argv", ["main.py", "--dataset", c["name"], "--data_path", tmp + "/", "--debug", "--batch_size", str(c["B"]),
                                      "--weight_size", c["ws"], "--epoch", "1", "--m_topk_rate", str(c["m_topk_rate"])])
    for m in ("dgl", "visdom"):                                   # import-only dependencies of main.py
        monkeypatch.setitem(sys.modules, m, types.ModuleType(m))
    monkeypatch.setattr(torch.cuda, "manual_seed_all", lambda s: None)
    monkeypatch.setattr(nn.Module, "cuda", lambda self, *a, **k: self)
    if not hasattr(np, "asfarray"):
        monkeypatch.setattr(np, "asfarray", lambda a, dtype=np.float64: np.asarray(a, dtype=dtype), raising=False)
    for name in [k for k in sys.modules if k == "main" or k == "Models" or k.startswith("utility")]:
        monkeypatch.delitem(sys.modules, name)
    # the drop-in file under the name the reference imports
    spec = importlib.util.spec_from_file_location("Models", os.path.join(ROOT, "mmssl_b200", "Models.py"))
    dropin = importlib.util.module_from_spec(spec)
    monkeypatch.setitem(sys.modules, "Models", dropin)
    spec.loader.exec_module(dropin)
    M = importlib.import_module("main")
    assert M.MMSSL is dropin.MMSSL and M.Discriminator is dropin.Discriminator
    assert dropin.args.embed_size == M.args.embed_size          # the reference's own flags reached the drop-in

    M.set_seed(c["seed"])
    tr = M.Trainer({})
    assert isinstance(tr.model, dropin.MMSSL)
    named = dict(tr.model.named_parameters())
    # seeded construction draws the reference's initial values (parameter creation order, Models.py:28-66) ...
    for k in LIVE:
        assert torch.equal(named[k].detach(), torch.from_numpy(z["G0/" + k])), k
    for k, v in tr.D.state_dict().items():
        assert torch.equal(v, torch.from_numpy(z["D0/" + k])), k
    # ... every random draw of the recorded run is replayed
    dg = M.data_generator
    dg.n_train = (c["steps"] - 1) * c["B"]
    tr.model.dropout = _Replay(z["mask_model"])
    tr.D.net[3], tr.D.net[7] = _Replay(z["mask_d1"]), _Replay(z["mask_d2"])
    draws = {"alpha": 0, "gumbel": 0, "sample": 0}

    def rand(*a, **k):
        t = torch.from_numpy(z["alpha"][draws["alpha"]]).clone()
        draws["alpha"] += 1
        assert tuple(t.shape) == tuple(a[0] if len(a) == 1 and not isinstance(a[0], int) else a)
        return t
    monkeypatch.setattr(torch, "rand", rand)

    def uniform_(self, *a, **k):
        self.copy_(torch.from_numpy(z["gumbel_u"][draws["gumbel"]]))
        draws["gumbel"] += 1
        return self
    monkeypatch.setattr(torch.Tensor, "uniform_", uniform_)

    def sample():
        s = z["sample"][draws["sample"]]
        draws["sample"] += 1
        return [list(map(int, s[0])), list(map(int, s[1])), list(map(int, s[2]))]
    monkeypatch.setattr(dg, "sample", sample)
    dummy = {k: np.ones(3) for k in ("recall", "precision", "ndcg", "hit_ratio")}
    tr.test = lambda users, is_val: dict(dummy, auc=0.)           # evaluation is not on this path (multiprocessing pool)

    got = {"Ggrad": [], "Gparam": [], "Dstate": []}
    tr.optimizer_D.register_step_pre_hook(lambda o, a, k: got["Ggrad"].append({n: named[n].grad.detach().clone() for n in LIVE}))
    tr.optimizer_D.register_step_post_hook(lambda o, a, k: got["Gparam"].append({n: named[n].detach().clone() for n in LIVE}))
    tr.optim_D.register_step_post_hook(lambda o, a, k: got["Dstate"].append({n: v.detach().clone() for n, v in tr.D.state_dict().items()}))
    tr.train()

    assert draws["sample"] == c["steps"] and len(got["Gparam"]) == c["steps"]
    for s in range(c["steps"]):
        for k in LIVE:
            e = rel_err(got["Ggrad"][s][k], torch.from_numpy(z["Ggrad/" + k][s]))
            assert e < 1e-4, (s, "Ggrad", k, e)
            e = rel_err(got["Gparam"][s][k], torch.from_numpy(z["Gparam/" + k][s]))
            assert e < 1e-4, (s, "Gparam", k, e)
        for k in cfg["d_state_names"]:
            # biases in front of a BatchNorm have an exactly-zero gradient in exact arithmetic: what Adam normalises there is
            # rounding noise of the reference's own torch ops (tests/fullstep_check.py treats them the same way)
            # (and the running means of those BatchNorms follow the noise-driven biases from the second iteration on)
            if "num_batches_tracked" in k or k in ("net.0.bias", "net.4.bias") or (s > 0 and k.endswith("running_mean")):
                continue
            # the Discriminator is the reference's own torch module on both sides; it sees our outputs (1e-6 away from the
            # reference's) and Adam on its scalar output bias shows 1.1e-4 after three iterations: 1e-3 for this bystander
            e = rel_err(got["Dstate"][s][k], torch.from_numpy(z["Dstate/" + k][s]))
            assert e < 1e-3, (s, "Dstate", k, e)
