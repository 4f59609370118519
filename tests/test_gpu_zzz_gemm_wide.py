"""General-width tcgen05 GEMM (csrc/gemm_wide.cu) against fp64, and the GAN side running on it.

NOT YET RUN ON A GPU: the kernel was written when round 1 had no GPU time left.  On the CPU it runs through the emulator's
functional model of the PTX it issues (tests/test_emu_tensor_core.py) -- protocol and indexing, not timing or true
asynchrony; it shares its descriptor / swizzle / pipeline helpers with proj_tc.cu, which is green on hardware, and
mbar_wait traps instead of hanging.  The file sorts last in the GPU suite.  gan_ops.GEMM_IMPL stays "simt" by default until
these tests have passed on a B200; the CUDA-graph capture test stays gated by MMSSL_RUN_UNVALIDATED=1."""
import os

import pytest
import torch

from tests.golden_util import rel_err

pytestmark = pytest.mark.gpu
_unvalidated = pytest.mark.skipif(os.environ.get("MMSSL_RUN_UNVALIDATED") != "1",
                                  reason="CUDA-graph capture of the full step has not run on a GPU yet (set MMSSL_RUN_UNVALIDATED=1)")


@pytest.mark.parametrize("m,n,k", [(64, 24, 96), (300, 200, 96), (2048, 1762, 7050), (1762, 7050, 2048), (2048, 7050, 1762),
                                   (2048, 881, 1762), (130, 257, 70), (5, 1, 8)])
def test_gemm_wide_vs_fp64(m, n, k):
    from mmssl_b200 import ops
    g = torch.Generator().manual_seed(m + n + k)
    a, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g)
    a_hi, a_lo = ops.split_bf16(a.cuda())
    b_hi, b_lo = ops.split_bf16(b.cuda())
    want = a.double() @ b.double().t()
    out = torch.full((m, n), float("nan"), device="cuda")
    ops.gemm_bf16x3_wide(a_hi, a_lo, b_hi, b_lo, m, n, k, out, alpha=0.5)
    assert rel_err(out, 0.5 * want) < 2e-5
    base = torch.randn(m, n, generator=g)
    out2 = base.clone().cuda()
    ops.gemm_bf16x3_wide(a_hi, a_lo, b_hi, b_lo, m, n, k, out2, alpha=-1.0, accumulate=True)
    assert rel_err(out2, base.double() - want) < 2e-5
    # a strided output (leading dimension > n, rows not 16-byte aligned): the scalar-store epilogue
    wide = torch.zeros(m, n + 3, device="cuda")
    ops.gemm_bf16x3_wide(a_hi, a_lo, b_hi, b_lo, m, n, k, wide[:, 1:n + 1], alpha=1.0)
    assert rel_err(wide[:, 1:n + 1], want) < 2e-5 and float(wide[:, 0].abs().max()) == 0 and float(wide[:, n + 1:].abs().max()) == 0


def test_gan_side_on_tensor_cores(monkeypatch):
    from mmssl_b200 import gan_ops
    from tests import fullstep_check, test_gpu_zz_gan as G
    monkeypatch.setattr(gan_ops, "GEMM_IMPL", "tc")
    G.test_d_step_on_gpu_matches_reference_trace()
    G.test_usim_and_real_rows(19445, 7050, 1024, 64)
    fullstep_check.run_and_check(dev="cuda", proj_impl="tc")


@_unvalidated
def test_full_step_cuda_graph_replay_equals_eager():
    """FullStep.capture(): the steady-state iteration as one CUDA graph == the same iterations run eagerly (same injected draws)."""
    from mmssl_b200.engine import LIVE
    from tests import fullstep_check
    z, c = fullstep_check.load_trace()
    runs = []
    for use_graph in (False, True):
        fs, P, t = fullstep_check.build(z, c, "cuda")
        torch.manual_seed(1)
        for s in range(3):                                   # the three recorded iterations bring the step into its steady state
            fs.step(*(t(z["sample"][s][j]) for j in range(3)))
        assert fs.steady()
        if use_graph:
            fs.capture()                                     # one warm-up iteration with its own draws + the capture
        else:
            torch.manual_seed(99)
            fs.step(*(t(z["sample"][0][j]) for j in range(3)))     # stands for the warm-up iteration; draws differ -> compare from a reset
        runs.append((fs, P))
    # identical starting point for the comparison: copy the eager twin's state into the graph twin
    (fe, Pe), (fg, Pg) = runs
    for k in LIVE:
        Pg[k].copy_(Pe[k]); fg.hs.m[k].copy_(fe.hs.m[k]); fg.hs.v[k].copy_(fe.hs.v[k])
    fg.hs.step_dev.copy_(fe.hs.step_dev)
    from mmssl_b200 import gan
    for k in gan.PARAMS + gan.BUFFERS:
        fg.D.t[k].copy_(fe.D.t[k])
    for k in gan.PARAMS:
        fg.D.m[k].copy_(fe.D.m[k]); fg.D.v[k].copy_(fe.D.v[k])
    fg.D.step_dev.copy_(fe.D.step_dev); fg.D.step = fe.D.step
    g = torch.Generator().manual_seed(5)
    B, I, d = c["B"], c["I"], c["d"]
    for s in range(3):
        mk = lambda n, w, p: ((torch.rand(n, w, generator=g) >= p) / (1 - p)).float().cuda()
        draws = dict(model_masks=[mk(I, d, 0.2) for _ in range(4)], d_masks1=[mk(2 * B, I // 4, 0.31) for _ in range(4)],
                     d_masks2=[mk(2 * B, I // 8, 0.5) for _ in range(4)], gumbel_u=torch.rand(B, I, generator=g).cuda(),
                     alpha=torch.rand(2 * B, generator=g).cuda())
        batch = [t(z["sample"][s][j]) for j in range(3)]
        oe, og = fe.step(*batch, **draws), fg.step(*batch, **draws)
        for k in ("batch_loss", "gp", "loss_D"):
            assert abs(float(oe[k]) - float(og[k])) <= 1e-5 * abs(float(oe[k])) + 1e-7, (s, k)
    for k in LIVE:
        assert rel_err(Pg[k], Pe[k]) < 1e-5, k
