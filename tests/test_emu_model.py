"""The whole hot path on the CPU: drop-in ``Models.MMSSL`` (autograd wrappers), the fused ``HotStep`` and three AdamW steps,
executed by the cuemu fiber emulator (tests/cuemu) against the golden vectors minted from the unmodified reference -- the
bodies of tests/test_gpu_model.py.  The tcgen05 GEMM is replaced by a host statement of its contract
(tests/cuemu/gemm_bf16x3_host.cpp: bf16 hi/lo operands, split-K partials), every other kernel is the real source."""
import pytest

from tests import test_gpu_model as M
from tests.cuemu import harness
from tests.golden_util import CASES, Golden, rel_err


@pytest.fixture
def emu(monkeypatch):
    harness.set_order("fwd")
    return harness.emulated_device(monkeypatch)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("proj_impl", ["tc", "simt"])
def test_model_forward_backward(emu, case, proj_impl):
    M.test_model_forward_backward_vs_reference(case, proj_impl)


@pytest.mark.parametrize("case", CASES)
def test_fused_hot_step(emu, case):
    M.test_fused_hot_step_vs_reference(case)


def test_three_adamw_steps_vs_oracle(emu):
    """tests/test_gpu_model.py::test_hot_step_graph_replay_and_adamw_vs_oracle without the CUDA graph."""
    from oracle import mmssl_oracle as O
    g = Golden("case_train_rand_k3")
    hs, P = M._hotstep(g, optimizer_step=True)
    cpu = O.CpuHotStep({k: v.clone() for k, v in g.params.items()}, g.image_feats, g.text_feats, g.graphs(), g.cfg["I"], g.oracle_cfg())
    got = [float(hs.run()[0]) for _ in range(3)]
    want = [cpu.step(g.users, g.pos, g.neg, dropout_masks=g.masks) for _ in range(3)]
    for a, b in zip(got, want):
        assert abs(a - b) < M.TOL * abs(b)
    for k in M.LIVE:
        assert rel_err(P[k], cpu.params[k]) < M.TOL, k
    assert int(hs.step_dev) == 3


@pytest.mark.parametrize("d,modal", [(128, "random"), (256, "random"), (128, "alias")])
def test_other_widths(emu, d, modal):
    M.test_hot_step_other_widths_vs_oracle(d, modal)
