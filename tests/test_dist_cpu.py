"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in mmssl_b200/parallel.py: the flat
gradient bucket of the data-parallel path and the row-sharded propagation schedule (partitioning +
one all-gather per half-layer, forward and backward).  The SpMM is injected: a CPU restatement here,
the CUDA operator on the box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmssl_b200 import parallel as par
from mmssl_b200.synthetic import csr_norm, make_bipartite

WORLD = 2
U, I, D, K = 203, 131, 16, 3


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _cpu_spmm(blk, x_full, *, c=None, alpha=1.0, epilogue=par.EPI_NONE, ysaved=None):
    v = torch.from_numpy(blk @ x_full.numpy())
    if c is not None:
        v = v + alpha * c
    if epilogue == par.EPI_SOFTMAX:
        v = torch.softmax(v, dim=-1)
    elif epilogue == par.EPI_SOFTMAX_BWD:
        v = ysaved * (v - (v * ysaved).sum(-1, keepdim=True))
    return v


def _cpu_softmax_bwd(y, g, alpha):
    g = alpha * g
    return y * (g - (g * y).sum(-1, keepdim=True))


def _problem():
    r = make_bipartite(U, I, 1500, seed=5)
    a_ui, a_iu = csr_norm(r).astype(np.float64), csr_norm(r.T.tocsr()).astype(np.float64)
    g = torch.Generator().manual_seed(1)
    u0, i0 = torch.randn(U, D, generator=g, dtype=torch.float64), torch.randn(I, D, generator=g, dtype=torch.float64)
    gu, gi = torch.randn(U, D, generator=g, dtype=torch.float64), torch.randn(I, D, generator=g, dtype=torch.float64)
    return a_ui, a_iu, u0, i0, gu, gi


def _reference(a_ui, a_iu, u0, i0, gu, gi):
    """single-process chain with autograd (same math as oracle.forward_closed's GCN loop)."""
    A = torch.from_numpy(a_ui.toarray()); B = torch.from_numpy(a_iu.toarray())
    u0 = u0.clone().requires_grad_(True); i0 = i0.clone().requires_grad_(True)
    s_u, s_i, i = u0, i0, i0
    for k in range(K):
        u = A @ i
        if k == K - 1:
            u = torch.softmax(u, -1)
        i = B @ u
        if k == K - 1:
            i = torch.softmax(i, -1)
        s_u = s_u + u; s_i = s_i + i
    ((s_u * gu).sum() + (s_i * gi).sum()).backward()
    return s_u.detach(), s_i.detach(), u0.grad, i0.grad


def _worker(rank, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        a_ui, a_iu, u0, i0, gu, gi = _problem()
        pu, pi = par.RowPartition(U, WORLD), par.RowPartition(I, WORLD)
        ops = {"ui": par.shard_rows_scipy(a_ui, pu, rank), "iu": par.shard_rows_scipy(a_iu, pi, rank),
               "iuT": par.shard_rows_scipy(a_iu.T.tocsr(), pu, rank), "uiT": par.shard_rows_scipy(a_ui.T.tocsr(), pi, rank)}
        gcn = par.RowShardedGCN(ops, pu, pi, K, _cpu_spmm, _cpu_softmax_bwd, rank)
        s_u, s_i, saved = gcn.forward(pu.local(u0, rank), pi.local(i0, rank))
        g_u0, g_i0 = gcn.backward(saved, pu.local(gu, rank), pi.local(gi, rank))
        full = [par.all_gather_rows(t, p) for t, p in ((s_u, pu), (s_i, pi), (g_u0, pu), (g_i0, pi))]
        # data-parallel bucket: mean of per-rank gradients
        grads = {"a": torch.full((5, 3), float(rank + 1)), "b": torch.arange(7, dtype=torch.float32) * (rank + 1)}
        b = par.GradBucket(grads)
        b.all_reduce_mean()
        if rank == 0:
            ret["full"] = [t.clone() for t in full]
            ret["gathers"] = gcn.n_gathers
            ret["bucket"] = {k: v.clone() for k, v in b.views.items()}
    finally:
        dist.destroy_process_group()


def test_row_sharded_gcn_and_bucket_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(_free_port(), ret), nprocs=WORLD, join=True)
    want = _reference(*_problem())
    for got, w in zip(ret["full"], want):
        assert torch.allclose(got, w, rtol=1e-10, atol=1e-12)
    # 1 (i_0) + 2 per layer - 1 (the last i is not needed) forward, 2 per layer backward
    assert ret["gathers"] == (1 + 2 * K - 1) + 2 * K
    assert torch.allclose(ret["bucket"]["a"], torch.full((5, 3), 1.5))
    assert torch.allclose(ret["bucket"]["b"], torch.arange(7, dtype=torch.float32) * 1.5)


def test_row_partition_padding():
    p = par.RowPartition(10, 4)
    assert p.block == 3 and [p.bounds(r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    x = torch.arange(20.).view(10, 2)
    assert torch.equal(p.local(x, 3)[0], x[9]) and float(p.local(x, 3)[1:].abs().sum()) == 0.0
