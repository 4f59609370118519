"""The two kernels that talk through NVSwitch multicast (`multimem.ld_reduce` / `multimem.st`), played for several ranks in ONE
process under the cuemu emulator: a fake multicast address range stands for n replicas (tests/cuemu/cuemu_ptx.cpp).
Both kernels are green on real multi-GPU boxes (tests/test_gpu_dist.py); this keeps their arithmetic and indexing under test
where there is no GPU at all.
  * mmssl_dp_fused_adamw  -- reduce-scatter of the gradient buckets + AdamW on the rank's slice + all-gather of the parameters;
  * mmssl_spmm_csr_f32 with y_mode = 1 -- the SpMM epilogue stores its row block into every rank's table."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from tests.cuemu import harness
from tests.golden_util import rel_err


@pytest.fixture
def emu(monkeypatch):
    harness.set_order("fwd")
    lib = harness.emulated_device(monkeypatch)
    raw = harness.emu_lib()
    raw.cuemu_mc_clear()
    yield raw
    raw.cuemu_mc_clear()


def _register(raw, replicas):
    """A dummy buffer's address range plays the multicast VA; returns (keepalive, multicast base address)."""
    nbytes = replicas[0].numel() * replicas[0].element_size()
    fake = torch.empty(nbytes, dtype=torch.uint8)
    ptrs = (C.c_void_p * len(replicas))(*[r.data_ptr() for r in replicas])
    raw.cuemu_mc_register(C.c_uint64(fake.data_ptr()), C.c_uint64(nbytes), len(replicas), ptrs)
    return fake, fake.data_ptr()


@pytest.mark.parametrize("world", [1, 3])
def test_fused_dp_optimiser_kernel(emu, world):
    from mmssl_b200 import _lib
    lib = _lib.load()
    n = 4 * 1000 * world                                  # flat bucket, slices of n / world floats
    g = torch.Generator().manual_seed(world)
    p0 = torch.randn(n, generator=g)
    params = [p0.clone() for _ in range(world)]           # one parameter bucket and one gradient bucket per "rank"
    grads = [torch.randn(n, generator=g) for _ in range(world)]
    keep_p, p_mc = _register(emu, params)
    keep_g, g_mc = _register(emu, grads)
    sl = n // world
    ms, vs = [torch.zeros(sl) for _ in range(world)], [torch.zeros(sl) for _ in range(world)]
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=5.5e-4)
    for step in (1, 2, 3):
        for gr in grads:
            gr.copy_(torch.randn(n, generator=g))
        ref.grad = torch.stack(grads).mean(0)
        opt.step()
        for r in range(world):                            # the ranks' kernels, one after the other (barriers are host-side)
            _lib.check(lib.mmssl_dp_fused_adamw(_lib.ptr(params[r]), C.c_void_p(p_mc), C.c_void_p(g_mc), _lib.ptr(ms[r]), _lib.ptr(vs[r]),
                                                r * sl, sl, 1.0 / world, step, 5.5e-4, 0.9, 0.999, 1e-8, 1e-2, None))
        for r in range(world):
            assert rel_err(params[r], ref) < 2e-6, (step, r)
            assert torch.equal(params[r], params[0])


def test_spmm_epilogue_multicast_store(emu):
    from mmssl_b200 import ops
    from mmssl_b200.graph import SparseOperand
    world, block, n_cols, d = 2, 40, 50, 64
    rng = np.random.default_rng(0)
    tables = [torch.zeros(world * block, d) for _ in range(world)]       # every rank's copy of the full [world*block, d] table
    keep, mc = _register(emu, tables)
    x = torch.randn(n_cols, d)
    want = torch.zeros(world * block, d, dtype=torch.float64)
    for rank in range(world):
        m = sp.random(block, n_cols, density=0.2, random_state=rank, dtype=np.float32).tocoo()
        op = SparseOperand(torch.from_numpy(m.row.astype(np.int64)), torch.from_numpy(m.col.astype(np.int64)), torch.from_numpy(m.data),
                           block, n_cols)
        op.tighten()
        local = tables[rank][rank * block:(rank + 1) * block]           # shape / strides only: the stores go through the multicast address
        ops.spmm(op, [x], [local], y_mode=1, y_raw=[mc + rank * block * d * 4])
        want[rank * block:(rank + 1) * block] = torch.from_numpy(m.tocsr() @ x.double().numpy())
    for t in tables:                                                      # both ranks hold both row blocks
        assert rel_err(t, want) < 2e-6


@pytest.mark.parametrize("y_mode", [1, 2])
def test_publish_rows_is_an_all_gather(emu, y_mode):
    """mmssl_publish_rows played for 3 ranks: afterwards every rank's table holds every rank's row block -- through the fake
    multicast address (mode 1) and through peer pointers (mode 2); strided source (a column half of a wider buffer)."""
    from mmssl_b200.rowshard_step import publish_rows
    world, block, d = 3, 17, 64
    g = torch.Generator().manual_seed(y_mode)
    tables = [torch.zeros(world * block, d) for _ in range(world)]
    keep, mc = _register(emu, tables)
    wide = [torch.randn(block, 2 * d, generator=g) for _ in range(world)]
    for rank in range(world):
        src = wide[rank][:, d:]                                            # ld = 2d
        local = tables[rank][rank * block:(rank + 1) * block]
        off = rank * block * d * 4
        if y_mode == 1:
            publish_rows(src, local, y_mode=1, y_raw=mc + off)
        else:
            publish_rows(src, local, y_mode=2, y_peers=[tables[r].data_ptr() + off for r in range(world) if r != rank])
    want = torch.cat([w[:, d:] for w in wide])
    for t in tables:
        assert torch.equal(t, want)


def test_multicast_all_reduce_kernel(emu):
    """mmssl_mc_allreduce_sum for 3 ranks: everyone reads the sum of the three copies through the (fake) multicast address."""
    from mmssl_b200 import _lib
    lib = _lib.load()
    world, n = 3, 4 * 333
    g = torch.Generator().manual_seed(9)
    copies = [torch.randn(n, generator=g) for _ in range(world)]
    keep, mc = _register(emu, copies)
    want = copies[0].double() + copies[1].double() + copies[2].double()
    for rank in range(world):
        out = torch.empty(n)
        _lib.check(lib.mmssl_mc_allreduce_sum(C.c_void_p(mc), _lib.ptr(out), n, None))
        assert rel_err(out, want) < 1e-6
