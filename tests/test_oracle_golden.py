"""Pin the CPU oracle against golden vectors minted from the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import mmssl_oracle as O
from tests.golden_util import CASES, Golden, rel_err

TOL = 2e-6  # oracle and reference are both torch-CPU fp32: only summation-order noise is allowed


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("variant", ["literal", "closed"])
def test_forward_losses_grads(case, variant):
    g = Golden(case)
    cfg = g.oracle_cfg()
    params = {k: v.clone().requires_grad_(True) for k, v in g.params.items()}
    fwd = O.forward_literal if variant == "literal" else O.forward_closed
    outs = fwd(params, g.image_feats, g.text_feats, g.graphs(), cfg,
               dropout_masks=g.masks if g.train else None, training=g.train)
    assert outs[0] is outs[6] and outs[1] is outs[7]
    for j in range(12):
        assert rel_err(outs[j], g.outs[j]) < TOL, (case, variant, j)
    total, parts = O.hot_loss(outs, g.users, g.pos, g.neg, g.cfg["I"], cfg, literal=(variant == "literal"))
    assert abs(float(parts["mf"]) - g.losses["mf"]) < 1e-6
    assert abs(float(parts["emb"]) - g.losses["emb"]) < 1e-9
    assert abs(float(parts["feat_reg"]) - g.losses["feat_reg"]) < 1e-9
    assert abs(float(parts["cl"]) - (g.losses["cl1"] + g.losses["cl2"])) < 1e-4 * abs(g.losses["cl1"] + g.losses["cl2"])
    assert abs(float(total) - g.losses["total"]) < 1e-6 * max(1.0, abs(g.losses["total"]))
    total.backward()
    for k, ref in g.grads.items():
        if k == "weight_dict.w_q":   # reference gets only rounding noise here
            got = params[k].grad
            assert got is None or float(got.abs().max()) < 1e-10
            continue
        got = params[k].grad
        assert got is not None, k
        assert rel_err(got, ref) < 5e-5, (case, variant, k, rel_err(got, ref))


@pytest.mark.parametrize("case", CASES)
def test_infonce_block_independent(case):
    g = Golden(case)
    cfg = g.oracle_cfg()
    z1, z2 = g.outs[8][g.users], g.outs[6][g.users]
    a = O.infonce(z1, z2, cfg, block=16)
    b = O.infonce_literal(z1, z2, cfg, block=16)
    assert abs(float(a) - g.losses["cl_small_block"]) < 1e-5
    assert abs(float(b) - g.losses["cl_small_block"]) < 1e-5
    assert abs(float(a) - g.losses["cl1"]) < 1e-5


def test_graph_normalisation_matches_golden():
    import numpy as np
    import scipy.sparse as sp
    g = Golden("case_eval_alias_k2")
    U, I = g.cfg["U"], g.cfg["I"]
    r = sp.csr_matrix((np.ones(len(g.z["train_rows"]), np.float32), (g.z["train_rows"], g.z["train_cols"])), shape=(U, I))
    ui, iu = O.build_graphs(r)
    gui, giu = g.graphs()[:2]
    assert rel_err(ui.to_dense(), gui.to_dense()) < 1e-7
    assert rel_err(iu.to_dense(), giu.to_dense()) < 1e-7


def test_adamw_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(37, 5)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=5.5e-4)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        gr = torch.randn_like(p)
        ref.grad = gr.clone()
        opt.step()
        O.adamw_step(p, gr, m, v, step, lr=5.5e-4)
        assert rel_err(p, ref) < 1e-6


# ------------------------------------------------------------------------------------------ evaluation path (section 8f row 3)
@pytest.mark.parametrize("case", ["eval_random", "eval_ties", "eval_short"])
@pytest.mark.parametrize("split", ["test", "val"])
def test_eval_oracle_matches_reference(case, split):
    """oracle/eval_oracle.py == the unmodified reference's test_torch / test_one_user / ranklist_by_heapq."""
    import os
    from oracle import eval_oracle as EO
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", case + ".npz"))
    Ks = [int(k) for k in g["Ks"]]
    out = EO.evaluate(g["ua"], g["ia"], g[f"{split}_users"], g["train_indptr"], g["train_indices"], g[f"{split}_indptr"],
                      g[f"{split}_indices"], Ks)
    assert np.array_equal(out["ranked"], g[f"{split}_ranked"])          # incl. the tie order
    assert np.array_equal(out["hits"], g[f"{split}_hits"])
    np.testing.assert_allclose(out["per_user"], g[f"{split}_per_user"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(out["result"], g[f"{split}_result"], rtol=0, atol=1e-14)


def _model_vs_oracle(ua, ia, users, tr_ptr, tr_idx, he_ptr, he_idx, Ks, seed=0):
    from oracle import eval_oracle as EO
    from tests.eval_kernel_model import rank_one_user
    ref = EO.evaluate(ua, ia, users, tr_ptr, tr_idx, he_ptr, he_idx, Ks)
    rating = EO.scores(ua, ia, users)
    rng = np.random.default_rng(seed)
    for n, u in enumerate(users):
        tr = np.sort(tr_idx[tr_ptr[u]:tr_ptr[u + 1]])
        he = np.sort(he_idx[he_ptr[u]:he_ptr[u + 1]])
        ranked, met = rank_one_user(rating[n], tr, he, Ks, rng)
        want = ref["ranked"][n]
        assert np.array_equal(ranked, want[want >= 0]), (n, u)
        np.testing.assert_allclose(met, ref["per_user"][n], rtol=0, atol=1e-12)


@pytest.mark.parametrize("case", ["eval_random", "eval_ties", "eval_short"])
def test_eval_kernel_algorithm_model_on_golden(case):
    """The selection algorithm of csrc/eval.cu (executable model) == oracle == reference, incl. ties and short lists."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", case + ".npz"))
    _model_vs_oracle(g["ua"], g["ia"], g["test_users"], g["train_indptr"], g["train_indices"], g["test_indptr"],
                     g["test_indices"], [int(k) for k in g["Ks"]])


def test_eval_kernel_algorithm_model_many_items_and_ties():
    """Several compactions per user (I = 3000 > buffer), heavy ties, negative / zero / signed-zero scores."""
    rng = np.random.default_rng(3)
    U, I, d = 12, 3000, 8
    ua = (np.round(rng.standard_normal((U, d)) * 2) / 2).astype(np.float32)
    ia = (np.round(rng.standard_normal((I, d)) * 2) / 2).astype(np.float32)
    ia[::7] = 0.0                                      # exact zeros; with negative user entries -> -0.0 products
    ua[3] = -np.abs(ua[3])
    tr_ptr = np.arange(0, (U + 1) * 40, 40, dtype=np.int64)
    tr_idx = np.concatenate([rng.choice(I, 40, replace=False) for _ in range(U)]).astype(np.int64)
    he_ptr = np.arange(0, (U + 1) * 25, 25, dtype=np.int64)
    he_idx = np.concatenate([rng.choice(I, 25, replace=False) for _ in range(U)]).astype(np.int64)
    _model_vs_oracle(ua, ia, np.arange(U), tr_ptr, tr_idx, he_ptr, he_idx, [1, 10, 20, 50, 64], seed=1)
