"""Full training iteration (mmssl_b200/fullstep.py: D step + G step + both optimisers + top-k graph rebuilds) on the GPU
against the trace recorded from the unmodified reference trainer, and the regraph kernels against torch / scipy.
The same bodies run on the CPU under the cuemu emulator (tests/test_emu_fullstep.py)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from tests import fullstep_check

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,w,k", [(32, 96, 4), (5, 7050, 1), (64, 1000, 17), (3, 40, 40), (7, 300, 0)])
def test_topk_rows_and_pairs(rows, w, k):
    from mmssl_b200.fullstep import pair_append, topk_rows
    g = torch.Generator().manual_seed(rows + w)
    x = torch.randn(rows, w, generator=g)
    x[:, ::7] = 0.0                                  # masked training items: exact ties at zero
    x[0, :5] = x[0, 5]                               # a run of equal values: the lower column wins
    ids = topk_rows(x.cuda(), k).cpu()
    assert ids.shape == (rows, k)
    if k:
        vals = torch.gather(x, 1, ids)
        want_vals = torch.topk(x, k, dim=-1).values
        assert torch.equal(vals, want_vals)                                     # same multiset of values, best first
        for r in range(rows):                                                   # ties: ascending column inside a run
            for j in range(k - 1):
                assert vals[r, j] > vals[r, j + 1] or ids[r, j] < ids[r, j + 1]
            assert len(set(ids[r].tolist())) == k
    users = torch.randperm(1000, generator=g)[:rows]
    px, py = pair_append(users.cuda(), ids.cuda())
    assert torch.equal(px.cpu(), users.repeat(1, k).view(-1)) and torch.equal(py.cpu(), ids.reshape(-1))   # main.py:398-399


@pytest.mark.parametrize("n_pairs", [0, 128, 5000])
def test_graphs_from_pairs_match_csr_norm(n_pairs):
    from mmssl_b200.fullstep import graphs_from_pairs
    from mmssl_b200.synthetic import csr_norm
    U, I = 120, 96
    rng = np.random.default_rng(n_pairs)
    x, y = rng.integers(0, U, n_pairs), rng.integers(0, I, n_pairs)             # with duplicates
    ui, iu = graphs_from_pairs(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), U, I)
    M = sp.csr_matrix((np.ones(n_pairs, np.float32), (x, y)), shape=(U, I))    # duplicates summed (main.py:379)
    for g, want in ((ui, csr_norm(M)), (iu, csr_norm(M.T.tocsr()))):
        got = g.fwd.to_scipy()
        got.sum_duplicates()
        assert g.nnz == n_pairs and (abs(got - want).max() < 1e-6 if n_pairs else got.nnz == 0)
        got_t = g.bwd.to_scipy()
        got_t.sum_duplicates()
        assert abs(got_t - want.T.tocsr()).max() < 1e-6 if n_pairs else got_t.nnz == 0


@pytest.mark.parametrize("proj_impl", ["tc", "simt"])
def test_full_step_matches_reference_trace(proj_impl):
    fs = fullstep_check.run_and_check(dev="cuda", proj_impl=proj_impl)
    assert fs.idx == 3 and fs.hs.graphs[2].nnz == 0 and fs.hs.graphs[4].nnz == 0


def test_full_step_own_random_draws_runs_and_learns():
    """Without injected draws (masks, Gumbel uniforms, interpolation weights from the CUDA generator): finite losses, the
    generator and the discriminator both move, the losses stay on the device."""
    z, c = fullstep_check.load_trace()
    fs, P, t = fullstep_check.build(z, c, "cuda")
    p0 = {k: v.clone() for k, v in P.items()}
    w0 = fs.D.t["net.0.weight"].clone()
    torch.manual_seed(0)
    for s in range(4):
        out = fs.step(*(t(z["sample"][s % 3][j]) for j in range(3)))
        for k in ("batch_loss", "G_lossf", "gp", "loss_D"):
            assert out[k].is_cuda and bool(torch.isfinite(out[k]).all()), k
    assert all(float((P[k] - p0[k]).abs().max()) > 0 for k in P if k in fs.hs.P)
    assert float((fs.D.t["net.0.weight"] - w0).abs().max()) > 0


@pytest.mark.parametrize("m_topk_rate,T", [(0.0, 1), (0.05, 2), (0.02, 3)])
def test_full_step_other_bookkeeping_regimes_vs_oracle(m_topk_rate, T):
    fullstep_check.regime_check("cuda", m_topk_rate, T, proj_impl="tc")


@pytest.mark.parametrize("d,I", [(128, 97), (256, 50)])
def test_full_step_other_shapes_vs_oracle(d, I):
    fullstep_check.random_problem_check("cuda", d=d, I=I, proj_impl="tc")
