"""The whole training iteration of the reference (D step, G step with the G_rate * G_lossf term, both optimisers, top-k
graph rebuilds) executed by the product's FullStep on the CPU under the cuemu emulator, against the trace recorded from
the unmodified reference trainer.  Same body as the GPU test (tests/fullstep_check.py)."""
import pytest

from tests import fullstep_check
from tests.cuemu import harness


@pytest.mark.parametrize("proj_impl", ["tc", "simt"])
def test_full_step_matches_reference_trace(monkeypatch, proj_impl):
    harness.set_order("fwd")
    harness.emulated_device(monkeypatch)
    fs = fullstep_check.run_and_check(dev="cpu", proj_impl=proj_impl)
    # T = 1: iteration 0 collected pairs, iteration 1 built graphs from them, iteration 2 rebuilt from empty lists
    assert fs.idx == 3 and fs.hs.graphs[2].nnz == 0 and fs.hs.graphs[4].nnz == 0
    # ... and from here on nothing changes any more: the iteration is a fixed sequence of launches (CUDA-graph capturable)
    assert fs.steady()
    g_before = fs.hs.graphs
    z, c = fullstep_check.load_trace()
    import numpy as np
    import torch
    t = lambda a: torch.from_numpy(np.asarray(a)).clone()
    out = fs.step(*(t(z["sample"][0][j]) for j in range(3)))
    assert fs.hs.graphs is g_before and fs.steady() and bool(torch.isfinite(out["batch_loss"]))


@pytest.fixture(params=["fwd", "rev"])
def emu(request, monkeypatch):
    harness.set_order(request.param)
    return harness.emulated_device(monkeypatch)


@pytest.mark.parametrize("rows,w,k", [(32, 96, 4), (5, 7050, 1), (64, 1000, 17), (3, 40, 40), (7, 300, 0)])
def test_topk_rows_and_pairs(emu, rows, w, k):
    from tests import test_gpu_zz_fullstep as G
    G.test_topk_rows_and_pairs(rows, w, k)


@pytest.mark.parametrize("n_pairs", [0, 128, 5000])
def test_graphs_from_pairs(emu, n_pairs):
    from tests import test_gpu_zz_fullstep as G
    G.test_graphs_from_pairs_match_csr_norm(n_pairs)


def test_own_random_draws(emu):
    from tests import test_gpu_zz_fullstep as G
    G.test_full_step_own_random_draws_runs_and_learns()


def test_full_step_with_tensor_core_gan_gemms(monkeypatch):
    from mmssl_b200 import gan_ops
    harness.set_order("fwd")
    harness.emulated_device(monkeypatch)
    monkeypatch.setattr(gan_ops, "GEMM_IMPL", "tc")
    fullstep_check.run_and_check(dev="cpu", proj_impl="tc")


@pytest.mark.parametrize("m_topk_rate,T", [(0.0, 1), (0.05, 2), (0.02, 3)])
def test_full_step_other_bookkeeping_regimes_vs_oracle(monkeypatch, m_topk_rate, T):
    harness.set_order("fwd")
    harness.emulated_device(monkeypatch)
    fullstep_check.regime_check("cpu", m_topk_rate, T)


@pytest.mark.parametrize("d,I", [(128, 97), (256, 50)])
def test_full_step_other_shapes_vs_oracle(monkeypatch, d, I):
    harness.set_order("fwd")
    harness.emulated_device(monkeypatch)
    fullstep_check.random_problem_check("cpu", d=d, I=I)


def test_steady_state_body_has_no_host_reads(monkeypatch):
    """CUDA-graph capturability of the steady-state iteration, as far as it can be checked without a GPU: no tensor is read back
    by the host (item / tolist / cpu / numpy / float() / int() / bool() / index) anywhere inside FullStep._body."""
    import torch
    from mmssl_b200 import gan_ops
    harness.set_order("fwd")
    harness.emulated_device(monkeypatch)
    monkeypatch.setattr(gan_ops, "GEMM_IMPL", "tc")
    z, c = fullstep_check.load_trace()
    fs, P, t = fullstep_check.build(z, c, "cpu", proj_impl="tc")
    for s in range(3):
        fs.step(*(t(z["sample"][s][j]) for j in range(3)))
    assert fs.steady()
    hits, state = [], {"on": False}

    def trap(name, orig):
        def f(self, *a, **k):
            if state["on"] and self.numel() >= 1:
                hits.append(name)
            return orig(self, *a, **k)
        return f
    for name in ("item", "tolist", "cpu", "numpy", "__float__", "__int__", "__bool__", "__index__"):
        monkeypatch.setattr(torch.Tensor, name, trap(name, getattr(torch.Tensor, name)))
    draws = fs._draws(None, None, None, None, None)
    fs.hs.set_indices(*(t(z["sample"][0][j]) for j in range(3)))
    state["on"] = True
    fs._body(*draws)
    state["on"] = False
    assert hits == []
