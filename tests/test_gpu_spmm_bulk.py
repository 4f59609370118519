"""Bulk-copy gather SpMM (csrc/spmm_bulk.cu, plan: mmssl_spmm_bulk_plan) against scipy fp64 and against the LDG kernel:
plain products, both ring sizes, several buckets per warp, all epilogues, row-indexed operands (alpha*C, running sums, saved
softmax output), rows cut at bucket boundaries (deterministic reduction) and heavy rows (vector reductions), empty rows, empty
graphs, strided operands, repeated launches (self-resetting counters / slots)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu

from tests.golden_util import rel_err  # noqa: E402
from tests.test_gpu_ops import _graph  # noqa: E402


def _bulk(nst=0, wpb=0, tpw=0, tma=0):
    from mmssl_b200 import ops
    return ops.SPMM_IMPL_BULK | nst | (wpb << 4) | (tpw << 8) | (ops.SPMM_BULK_TMA if tma else 0)


def test_bulk_plan_covers_every_nonzero_once():
    """Every position belongs to exactly one item, items of a bucket start inside it, long rows are cut at the boundaries."""
    g, ref = _graph(700, 500, 30000, seed=5, heavy_rows=2)
    b = g.fwd.bulk_plan()
    torch.cuda.synchronize()
    n_items, n_split, n_segs = b["totals"].cpu().tolist()
    items = b["items"].cpu().numpy().reshape(-1, 4)[:n_items]
    rowptr = g.fwd.rowptr.cpu().numpy()
    cover = np.zeros(g.fwd.nnz, np.int32)
    for row, lo, hi, split in items:
        cover[lo:hi] += 1
        assert rowptr[row] <= lo <= hi <= rowptr[row + 1]
        if split < 0:
            assert (lo, hi) == (rowptr[row], rowptr[row + 1]) and hi - lo <= 32
        else:
            assert hi - lo <= 32 and lo // 32 == (hi - 1) // 32
    assert (cover == 1).all()
    assert (np.diff(items[:, 1]) >= 0).all() and n_split > 0
    bk = b["buckets"].cpu().numpy().reshape(-1, 8)
    assert bk.shape[0] == g.fwd.nnz // 32 + 1
    seen = 0
    for t, (i0, n, row0, n_rows, nz0, nz1, _, _) in enumerate(bk):
        if n == 0:
            continue
        assert i0 == seen
        seen += n
        its = items[i0:i0 + n]
        assert (its[:, 1] // 32 == t).all() or (its[:, 1] == its[:, 2]).all() or ((its[:, 1] >= 32 * t) & (its[:, 1] < 32 * t + 32)).all()
        assert nz0 == its[0, 1] and nz1 == its[-1, 2] and nz1 - 32 * t <= 64
        assert row0 == its[0, 0] and n_rows == its[-1, 0] - row0 + 1
    assert seen == n_items


@pytest.mark.parametrize("d", [64, 128, 256])
@pytest.mark.parametrize("nrhs", [1, 2])
@pytest.mark.parametrize("variant", [(0, 0, 0, 0), (2, 2, 3, 0), (4, 4, 1, 1), (2, 8, 1, 0), (0, 0, 0, 1)])
def test_spmm_bulk_plain(d, nrhs, variant):
    from mmssl_b200 import ops
    g, ref = _graph(700, 500, 30000, seed=d + nrhs, heavy_rows=2)
    torch.manual_seed(0)
    wide = torch.randn(500, nrhs * d + 8, device="cuda")
    xs = [wide[:, r * d:(r + 1) * d] for r in range(nrhs)]          # strided views (ld != d)
    impl = _bulk(*variant)
    ys = ops.spmm(g.fwd, xs, impl=impl)
    for x, y in zip(xs, ys):
        assert rel_err(y, torch.from_numpy(ref @ x.double().cpu().numpy())) < 2e-6
    xt = torch.randn(700, d, device="cuda")
    yt = ops.spmm(g.bwd, [xt], impl=impl)[0]
    assert rel_err(yt, torch.from_numpy(ref.T @ xt.double().cpu().numpy())) < 2e-6
    yt2 = ops.spmm(g.bwd, [xt], impl=impl)[0]                       # counters / heavy slots reset themselves
    assert rel_err(yt2, yt) < 1e-6
    g2, _ = _graph(600, 500, 20000, seed=d, heavy_rows=40)          # ~170 nnz in each of 40 rows: cut, reduced in bucket order
    y_a = ops.spmm(g2.fwd, [xs[0]], impl=impl)[0]
    y_b = ops.spmm(g2.fwd, [xs[0]], impl=impl)[0]
    assert torch.equal(y_a, y_b)


@pytest.mark.parametrize("d", [64, 128, 256])
@pytest.mark.parametrize("nst", [2, 4])
@pytest.mark.parametrize("tma", [0, 1])
def test_spmm_bulk_epilogues(d, nst, tma):
    from mmssl_b200 import ops
    impl = _bulk(nst, tma=tma)
    g, ref = _graph(300, 260, 9000, seed=7 + d, heavy_rows=1)
    torch.manual_seed(1)
    x = torch.randn(260, d, device="cuda")
    c = torch.randn(300, d, device="cuda")
    base = torch.from_numpy(ref @ x.double().cpu().numpy())
    v = base + 0.25 * c.double().cpu()
    sb = torch.randn(300, d, device="cuda")
    s = torch.empty(300, d, device="cuda")
    y = ops.spmm(g.fwd, [x], cs=[c], alpha=0.25, epilogue=ops.EPI_SOFTMAX, ss=[s], s_mode=2, sbases=[sb], impl=impl)[0]
    want = torch.softmax(v, dim=-1)
    assert rel_err(y, want) < 5e-6
    assert rel_err(s, sb.double().cpu() + want) < 5e-6
    ops.spmm(g.fwd, [x], cs=[c], alpha=0.25, epilogue=ops.EPI_NONE, ss=[s], s_mode=1, impl=impl)
    assert rel_err(s, sb.double().cpu() + want + v) < 5e-6
    ysv = torch.softmax(torch.randn(300, d, device="cuda"), -1)
    t = ops.spmm(g.fwd, [x], cs=[c], alpha=0.25, epilogue=ops.EPI_SOFTMAX_BWD, ysaved=[ysv], impl=impl)[0]
    yd = ysv.double().cpu()
    assert rel_err(t, yd * (v - (v * yd).sum(-1, keepdim=True))) < 5e-6
    acc = c.clone()                                                 # in place: y = y + A x (C aliases Y)
    ops.spmm(g.fwd, [x], [acc], cs=[acc], alpha=1.0, impl=impl)
    assert rel_err(acc, base + c.double().cpu()) < 5e-6
    # two right-hand sides with the operands as halves of one buffer (engine.U2 / I2 layout)
    x2 = torch.randn(260, 2 * d, device="cuda")
    c2 = torch.randn(300, 2 * d, device="cuda")
    y2 = torch.empty(300, 2 * d, device="cuda")
    ops.spmm(g.fwd, [x2[:, :d], x2[:, d:]], [y2[:, :d], y2[:, d:]], cs=[c2[:, :d], c2[:, d:]], alpha=2.0, impl=impl)
    assert rel_err(y2, torch.from_numpy(ref @ x2.double().cpu().numpy()) + 2.0 * c2.double().cpu()) < 5e-6


def test_spmm_bulk_many_short_and_empty_rows():
    """More than 32 items and more than 16 rows per bucket (rows of 0-2 non-zeros): item batches and operand groups roll over."""
    from mmssl_b200 import ops
    from mmssl_b200.graph import BipartiteGraph
    rng = np.random.default_rng(3)
    n_rows, n_cols = 4000, 300
    deg = rng.integers(0, 3, n_rows)
    deg[100:180] = 0                                                # 80 consecutive empty rows
    deg[2000] = 45; deg[2001] = 33; deg[2002] = 32; deg[2003] = 31  # around the cut threshold
    r = np.repeat(np.arange(n_rows), deg)
    c = rng.integers(0, n_cols, len(r))
    v = rng.standard_normal(len(r)).astype(np.float32)
    ref = sp.coo_matrix((v.astype(np.float64), (r, c)), shape=(n_rows, n_cols)).tocsr()
    g = BipartiteGraph(torch.from_numpy(r).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(v).cuda(), (n_rows, n_cols))
    x = torch.randn(n_cols, 64, device="cuda")
    cc = torch.randn(n_rows, 64, device="cuda")
    s = torch.randn(n_rows, 64, device="cuda")
    s0 = s.clone()
    for nst, tma in ((2, 0), (4, 0), (2, 1), (4, 1)):
        s.copy_(s0)
        y = ops.spmm(g.fwd, [x], cs=[cc], alpha=-1.5, ss=[s], s_mode=1, impl=_bulk(nst, tma=tma))[0]
        want = torch.from_numpy(ref @ x.double().cpu().numpy()) - 1.5 * cc.double().cpu()
        assert rel_err(y, want) < 2e-6
        assert rel_err(s, s0.double().cpu() + want) < 2e-6
    y = ops.spmm(g.fwd, [x], epilogue=ops.EPI_SOFTMAX, impl=_bulk())[0]
    assert rel_err(y, torch.softmax(torch.from_numpy(ref @ x.double().cpu().numpy()), -1)) < 5e-6


def test_spmm_bulk_empty_graph():
    from mmssl_b200 import ops
    from mmssl_b200.graph import BipartiteGraph
    e = torch.zeros(0, dtype=torch.int64, device="cuda")
    g = BipartiteGraph(e, e, torch.zeros(0, device="cuda"), (40, 30))
    x = torch.randn(30, 64, device="cuda")
    assert float(ops.spmm(g.fwd, [x], impl=_bulk())[0].abs().max()) == 0.0
    y = ops.spmm(g.fwd, [x], epilogue=ops.EPI_SOFTMAX, impl=_bulk())[0]
    assert rel_err(y, torch.full((40, 64), 1 / 64)) < 1e-6


def test_spmm_bulk_heavy_rows_and_zipf_columns():
    from mmssl_b200 import ops
    from mmssl_b200.graph import BipartiteGraph
    rng = np.random.default_rng(9)
    n_rows, n_cols, nnz = 3000, 5000, 90000
    pw = 1.0 / np.arange(1, n_cols + 1); pw /= pw.sum()
    r = rng.integers(0, n_rows, nnz); r[:20000] = rng.integers(0, 3, 20000)     # three rows of ~6.7k non-zeros: heavy mode
    c = rng.choice(n_cols, nnz, p=pw)
    v = rng.standard_normal(nnz).astype(np.float32)
    ref = sp.coo_matrix((v.astype(np.float64), (r, c)), shape=(n_rows, n_cols)).tocsr()
    g = BipartiteGraph(torch.from_numpy(r).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(v).cuda(), (n_rows, n_cols))
    for d in (64, 128):
        x = torch.randn(n_cols, d, device="cuda")
        cs = torch.randn(n_rows, d, device="cuda")
        for _ in range(2):
            y = ops.spmm(g.fwd, [x], cs=[cs], alpha=0.5, epilogue=ops.EPI_SOFTMAX, impl=_bulk(tma=_))[0]
            want = torch.softmax(torch.from_numpy(ref @ x.double().cpu().numpy()) + 0.5 * cs.double().cpu(), -1)
            assert rel_err(y, want) < 1e-5
        yt = ops.spmm(g.bwd, [torch.ones(n_rows, d, device="cuda")], impl=_bulk())[0]
        assert rel_err(yt, torch.from_numpy(np.asarray(ref.T.sum(1))).expand(-1, d)) < 1e-5
