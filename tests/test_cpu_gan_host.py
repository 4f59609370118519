"""The GAN-side orchestration of the product (mmssl_b200/gan.py: which op runs when, on what) replayed against the trace
recorded from the unmodified reference trainer, with the device ops injected as their torch-CPU specification
(tests/gan_ops_cpu.py).  What stays for the GPU suite is each CUDA op against the function of the same name."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from mmssl_b200 import gan
from oracle import gan_oracle as GO
from tests import gan_ops_cpu as K
from tests.golden_util import rel_err

DEAD_BIAS = {"net.0.bias": "net.0.weight", "net.4.bias": "net.4.weight"}     # see tests/test_gan_oracle.py


@pytest.fixture(scope="module")
def trace():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "gan_trace.npz"))
    c = json.loads(str(z["cfg"]))
    R = sp.csr_matrix((np.ones(len(z["train_rows"]), np.float32), (z["train_rows"], z["train_cols"])), shape=(c["U"], c["I"]))
    R.sort_indices()
    return z, c, torch.from_numpy(R.indptr.astype(np.int64)), torch.from_numpy(R.indices.astype(np.int64)), R


def test_d_step_sequence_matches_reference_trace(trace):
    z, c, indptr, indices, R = trace
    t = lambda a: torch.from_numpy(np.asarray(a)).clone()
    D = gan.DiscriminatorState({k[3:]: t(z[k]) for k in z.files if k.startswith("D0/")})
    hp = gan.GanHyper(gp_rate=c["gp_rate"], G_rate=c["G_rate"], D_lr=c["D_lr"], log_log_scale=c["log_log_scale"],
                      real_data_tau=c["real_data_tau"], ui_pre_scale=c["ui_pre_scale"])
    for s in range(c["steps"]):
        users = t(z["sample"][s][0])
        ui, img, txt = (t(z["u_sim"][5 * s + j]) for j in range(3))
        m1 = [t(z["mask_d1"][4 * s + j]) for j in range(4)]
        m2 = [t(z["mask_d2"][4 * s + j]) for j in range(4)]
        out = gan.d_step(K, D, hp, img, txt, ui, users, indptr, indices, t(z["gumbel_u"][s]), t(z["alpha"][s]).view(-1), m1, m2)
        n = out["n"]
        assert abs(float(out["gp"]) - float(z["gp"][s])) <= 1e-4 * abs(float(z["gp"][s]))
        assert abs(100 * float(out["lossf_sum"]) / n - float(z["D_out"][4 * s].mean())) < 1e-3
        assert abs(100 * float(out["lossr_sum"]) / n - float(z["D_out"][4 * s + 1].mean())) < 1e-3
        for k in gan.PARAMS:
            want = t(z["Dgrad/" + k][s])
            if k in DEAD_BIAS:
                assert float(out["grads"][k].abs().max()) < 1e-5 * float(np.abs(z["Dgrad/" + DEAD_BIAS[k]][s]).max())
            else:
                assert rel_err(out["grads"][k].view_as(want), want) < 2e-4, (s, k)
        for k in gan.PARAMS + gan.BUFFERS:
            want = t(z["Dstate/" + k][s])
            if k in DEAD_BIAS:
                continue
            if k.endswith("running_mean"):
                assert float((D.t[k] - want).abs().max()) <= 1.01 * c["D_lr"] * (s + 1), (s, k)
            elif want.dtype == torch.int64:
                assert int(D.t[k]) == int(want)
            else:
                assert rel_err(D.t[k], want) < 2e-4, (s, k)
        # the G step's D call moves the BatchNorm buffers once more before the next D step
        gan.d_forward(K, D, t(z["D_in"][4 * s + 3]), m1[3], m2[3])


def test_g_side_input_gradient_matches_autograd(trace):
    z, c, indptr, indices, R = trace
    t = lambda a: torch.from_numpy(np.asarray(a)).clone()
    state = {k[3:]: t(z[k]) for k in z.files if k.startswith("D0/")}
    hp = gan.GanHyper(G_rate=c["G_rate"])
    x = t(z["D_in"][3]).requires_grad_(True)
    m1, m2 = t(z["mask_d1"][3]), t(z["mask_d2"][3])
    S = {k: v.clone() for k, v in state.items()}
    loss = -hp.G_rate * GO.discriminator(x, S, m1, m2).mean()
    want = torch.autograd.grad(loss, x)[0]
    B = x.shape[0] // 2
    ci, ct = {"y": x.detach()[:B]}, {"y": x.detach()[B:]}
    D = gan.DiscriminatorState({k: v.clone() for k, v in state.items()})
    s_sum, gi, gt = gan.g_side(K, D, hp, ci, ct, m1, m2)
    assert rel_err(torch.cat((gi, gt)), want) < 1e-4
    assert abs(-100 * float(s_sum) / (2 * B) - float(loss) / hp.G_rate) < 1e-3
    for k in ("net.2.running_mean", "net.6.running_var"):
        assert rel_err(D.t[k], S[k]) < 1e-5                                   # buffers advanced like nn.BatchNorm1d


def test_u_sim_forward_backward_sequence(trace):
    z, c, indptr, indices, R = trace
    g = torch.Generator().manual_seed(0)
    U, I, d = c["U"], c["I"], 16
    uf = torch.randn(U, d, generator=g, requires_grad=True)
    itf = torch.randn(I, d, generator=g, requires_grad=True)
    users = torch.from_numpy(z["sample"][0][0]).clone()
    go = torch.randn(len(users), I, generator=g)
    want_y = GO.u_sim(users.tolist(), uf, itf, R, batch_size=c["B"])
    wu, wi = torch.autograd.grad((want_y * go).sum(), [uf, itf])
    cache = gan.u_sim_forward(K, uf.detach(), itf.detach(), users, indptr, indices)
    assert rel_err(cache["y"], want_y) < 1e-5
    gu, gi = torch.zeros(U, d), torch.zeros(I, d)
    gan.u_sim_backward(K, cache, go, itf.detach(), indptr, indices, gu, gi)
    assert rel_err(gu, wu) < 1e-4 and rel_err(gi, wi) < 1e-4
