"""The GAN-side kernels (csrc/gan.cu) executed on the CPU by the cuemu fiber emulator (tests/cuemu), through the
product's own Python binding, against their specification -- the bodies of tests/test_gpu_zz_gan.py, unchanged.
This is what stands in for a GPU run while the container has none: barriers, shuffles, indexing, reduction order and
the ctypes marshalling are the real ones; see tests/cuemu/include/cuemu.h for what the emulator does not cover."""
import pytest

from tests import test_gpu_zz_gan as G
from tests.cuemu import harness


@pytest.fixture(params=["fwd", "rev"])
def emu(request, monkeypatch):
    harness.set_order(request.param)
    return harness.emulated_device(monkeypatch)


@pytest.mark.parametrize("n,h", [(64, 24), (50, 33), (130, 70)])
def test_bn_ops(emu, n, h):
    G.test_bn_ops(n, h)


@pytest.mark.parametrize("n,h", [(64, 12), (37, 5), (300, 40)])
def test_head_ops(emu, n, h):
    G.test_head_ops(n, h)


@pytest.mark.parametrize("n,w", [(64, 96), (9, 700)])
def test_gp_rows_interpolate_and_axpy(emu, n, w):
    G.test_gp_rows_interpolate_and_axpy(n, w)


def test_usim_and_real_rows(emu):
    G.test_usim_and_real_rows(120, 96, 32, 64)


def test_d_step_matches_reference_trace(emu):
    G.test_d_step_on_gpu_matches_reference_trace()


@pytest.fixture
def tc_gemm(monkeypatch):
    from mmssl_b200 import gan_ops
    monkeypatch.setattr(gan_ops, "GEMM_IMPL", "tc")


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_tensor_core_gemm_route_operand_layouts(emu, tc_gemm, ta, tb):
    """gan_ops.mm through the bf16 hi/lo splits and the wide-GEMM entry point (contract stand-in on the CPU): the four
    transpose combinations, ragged sizes (k, n not multiples of 8 / of the tiles), the alpha epilogue."""
    import torch
    from mmssl_b200 import gan_ops as K
    from tests.golden_util import rel_err
    g = torch.Generator().manual_seed(3)
    m, n, k = 37, 45, 77
    a = torch.randn((k, m) if ta else (m, k), generator=g)
    b = torch.randn((n, k) if tb else (k, n), generator=g)
    got = K.mm(a, b, ta=ta, tb=tb, alpha=-0.25)
    A = a.double().t() if ta else a.double()
    B = b.double().t() if tb else b.double()
    assert rel_err(got, -0.25 * A @ B) < 2e-5


def test_d_step_trace_with_tensor_core_gemms(emu, tc_gemm):
    G.test_d_step_on_gpu_matches_reference_trace()
    G.test_usim_and_real_rows(120, 96, 32, 64)


def test_d_step_trace_with_library_gemms(emu, monkeypatch):
    """GEMM_IMPL = "cublas": the comparison route through torch.mm / addmm (the vendor library on a GPU)."""
    from mmssl_b200 import gan_ops
    monkeypatch.setattr(gan_ops, "GEMM_IMPL", "cublas")
    G.test_d_step_on_gpu_matches_reference_trace()


def test_weight_split_cache(emu, tc_gemm, monkeypatch):
    """tc route: the bf16 splits of registered weights are computed once per optimiser step, activations every call."""
    import torch
    from mmssl_b200 import gan_ops as K, ops
    from tests.golden_util import rel_err
    calls = {"n": 0}
    real = ops.split_bf16

    def counting(x, *a, **k):
        calls["n"] += 1
        return real(x, *a, **k)
    monkeypatch.setattr(ops, "split_bf16", counting)
    g = torch.Generator().manual_seed(0)
    w, x = torch.randn(40, 72, generator=g), torch.randn(24, 72, generator=g)
    K.weights_changed()
    K.register_weights([w])
    y1 = K.mm(x, w, tb=True)
    n1 = calls["n"]                                  # x and w split
    y2 = K.mm(x, w, tb=True)
    assert calls["n"] == n1 + 1                      # only the activation again
    assert rel_err(y1, x.double() @ w.double().t()) < 2e-5 and torch.equal(y1, y2)
    w.mul_(2.0)                                      # a torch-side in-place change is seen through the version counter
    y3 = K.mm(x, w, tb=True)
    assert rel_err(y3, x.double() @ w.double().t()) < 2e-5 and calls["n"] == n1 + 3
    K.weights_changed()                              # what adam() calls after the kernel has rewritten the weights
    K.mm(x, w, tb=True)
    assert calls["n"] == n1 + 5
