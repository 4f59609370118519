"""oracle/gan_oracle.py (GAN side + full training step) replayed against the trace recorded from three
steps of the unmodified reference `Trainer.train()` (tests/golden/gan_trace.npz, minted by make_golden_gan.py)."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import gan_oracle as GO
from oracle import mmssl_oracle as O
from tests.golden_util import rel_err

TOL = 2e-5      # same torch CPU ops in the same order; slack only for BLAS blocking / accumulation order
# A bias in front of a training-mode BatchNorm has an exactly zero gradient; what autograd returns is cancellation noise
# (1e-2 against weight gradients of 1e4) that Adam then normalises into a +-lr random walk of a parameter no output depends
# on.  Gradients: absolute tolerance at the scale of the layer; state: not compared.
DEAD_BIAS = {"net.0.bias": "net.0.weight", "net.4.bias": "net.4.weight"}


@pytest.fixture(scope="module")
def replay():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "gan_trace.npz"))
    c = json.loads(str(z["cfg"]))
    t = lambda a: torch.from_numpy(np.asarray(a))
    P = {k[3:]: t(z[k]) for k in z.files if k.startswith("G0/")}
    S = {k[3:]: t(z[k]).clone() for k in z.files if k.startswith("D0/")}
    R = sp.csr_matrix((np.ones(len(z["train_rows"]), np.float32), (z["train_rows"], z["train_cols"])), shape=(c["U"], c["I"]))
    cfg = O.HotPathConfig(embed_size=c["d"], n_layers=c["n_layers"], head_num=c["head_num"], id_cat_rate=c["id_cat_rate"],
                          model_cat_rate=c["model_cat_rate"], drop_rate=c["drop_rate"], tau=c["tau"], cl_rate=c["cl_rate"],
                          emb_decay=c["emb_decay"], feat_reg_decay=c["feat_reg_decay"], batch_size=c["B"], lr=c["lr"])
    g = GO.GanConfig(G_drop1=c["G_drop1"], G_drop2=c["G_drop2"], gp_rate=c["gp_rate"], G_rate=c["G_rate"], D_lr=c["D_lr"],
                     log_log_scale=c["log_log_scale"], real_data_tau=c["real_data_tau"], ui_pre_scale=c["ui_pre_scale"],
                     m_topk_rate=c["m_topk_rate"], T=c["T"])
    fs = GO.FullStep(P, S, t(z["image_feats"]), t(z["text_feats"]), R, cfg, g)
    traces = []
    for s in range(c["steps"]):
        users, pos, neg = (z["sample"][s][j] for j in range(3))
        traces.append(fs.step(users, pos, neg, [t(z["mask_model"][4 * s + j]) for j in range(4)],
                              [t(z["mask_d1"][4 * s + j]) for j in range(4)], [t(z["mask_d2"][4 * s + j]) for j in range(4)],
                              t(z["gumbel_u"][s]), t(z["alpha"][s])))
    return z, c, traces


def test_u_sim_and_discriminator_calls(replay):
    z, c, traces = replay
    for s, tr in enumerate(traces):
        for j in range(5):
            assert rel_err(tr["u_sim"][j], torch.from_numpy(z["u_sim"][5 * s + j])) < TOL, (s, j)
        for j in range(4):
            assert rel_err(tr["D_in"][j], torch.from_numpy(z["D_in"][4 * s + j])) < TOL, (s, j)
            assert rel_err(tr["D_out"][j], torch.from_numpy(z["D_out"][4 * s + j])) < TOL, (s, j)


def test_gradient_penalty_and_d_step(replay):
    z, c, traces = replay
    for s, tr in enumerate(traces):
        assert abs(float(tr["gp"]) - float(z["gp"][s])) <= TOL * abs(float(z["gp"][s])), s
        for k in GO.D_PARAMS:
            if k in DEAD_BIAS:
                scale = float(np.abs(z["Dgrad/" + DEAD_BIAS[k]][s]).max())
                assert float(tr["Dgrad"][k].abs().max()) < 1e-5 * scale and float(np.abs(z["Dgrad/" + k][s]).max()) < 1e-5 * scale
            else:
                assert rel_err(tr["Dgrad"][k], torch.from_numpy(z["Dgrad/" + k][s])) < 1e-4, (s, k)
        for k in c["d_state_names"]:
            if k in DEAD_BIAS:
                continue
            want = torch.from_numpy(np.asarray(z["Dstate/" + k][s]))
            if k.endswith("running_mean"):      # carries the dead bias of the Linear in front: bounded by its +-lr walk
                assert float((tr["Dstate"][k] - want).abs().max()) <= 1.01 * c["D_lr"] * (s + 1), (s, k)
            elif want.dtype == torch.int64:
                assert int(tr["Dstate"][k]) == int(want)          # BatchNorm num_batches_tracked: 3 calls before the step
            else:
                assert rel_err(tr["Dstate"][k], want) < 1e-4, (s, k)


def test_g_step_gradients_and_parameters(replay):
    z, c, traces = replay
    for s, tr in enumerate(traces):
        for k, gr in tr["Ggrad"].items():
            assert rel_err(gr, torch.from_numpy(z["Ggrad/" + k][s])) < 1e-4, (s, k)
        for k, p in tr["Gparam"].items():
            assert rel_err(p, torch.from_numpy(z["Gparam/" + k][s])) < 1e-4, (s, k)


def test_modality_graph_rebuilds(replay):
    """Step 1 rebuilds from the (users tiled, ids flattened) pairs of step 0; step 2 rebuilds empty graphs."""
    z, c, traces = replay
    assert "graphs" not in traces[0] and int(z["graph_idx/n"]) == 8
    for s in (1, 2):
        for j, gph in enumerate(traces[s]["graphs"]):
            n = 4 * (s - 1) + j
            gc = gph.coalesce()
            assert tuple(gc.shape) == tuple(int(v) for v in z["graph_shape"][n])
            assert np.array_equal(gc.indices().numpy(), z[f"graph_idx/{n}"])
            np.testing.assert_allclose(gc.values().numpy(), z[f"graph_val/{n}"], rtol=1e-6)
    assert traces[1]["graphs"][0]._nnz() > 0 and traces[2]["graphs"][0]._nnz() == 0
