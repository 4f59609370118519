"""Staged-gather SpMM (csrc/spmm_bulk.cu, plan: mmssl_spmm_bulk_plan) against scipy fp64: the bucket plan's invariants, plain
products with both copy engines (warp-wide cp.async / one TMA bulk copy per row), several buckets per warp, all epilogues,
row-indexed operands (alpha*C, running sums, saved softmax output), long rows cut into chunks (deterministic reduction) and
heavy rows (vector reductions), empty rows, empty graphs, strided operands, repeated launches (self-resetting counters / slots)."""
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


def _check_plan(op):
    """Every position belongs to exactly one bucket, buckets hold <= 32 positions and <= 8 whole rows (or one chunk of a long row),
    rows of <= 32 non-zeros are never cut, every row is covered, buckets come in row order."""
    b = op.bulk_plan()
    n = b["n_buckets"]
    bk = b["buckets"].cpu().numpy().reshape(-1, 8)[:n]
    rowptr = op.rowptr.cpu().numpy()
    n_rows = op.n_rows
    cover = np.zeros(max(op.nnz, 1), np.int32)
    row_seen = np.zeros(n_rows, np.int32)
    st = b["split_table"].cpu().numpy().reshape(-1, 4)
    last_row = -1
    for row0, nr, nz0, cnt, split, seg, _, _ in bk:
        assert 1 <= nr <= 8 and 0 <= cnt <= 32
        cover[nz0:nz0 + cnt] += 1
        if split < 0:
            assert row0 > last_row
            assert nz0 == rowptr[row0] and nz0 + cnt == rowptr[row0 + nr]
            assert (np.diff(rowptr[row0:row0 + nr + 1]) <= 32).all()
            row_seen[row0:row0 + nr] += 1
            last_row = row0 + nr - 1
        else:
            assert nr == 1 and rowptr[row0 + 1] - rowptr[row0] > 32
            chunks = -(-(rowptr[row0 + 1] - rowptr[row0]) // 32)
            assert st[split, 1] == chunks and 0 <= seg < chunks and nz0 == rowptr[row0] + 32 * seg
            assert st[split, 3] == (1 if chunks > 32 else 0)
            if seg == 0:
                row_seen[row0] += 1
                assert row0 > last_row
                last_row = row0
    if op.nnz:
        assert (cover[:op.nnz] == 1).all()
    assert (row_seen == 1).all()
    return bk


def test_bulk_plan_covers_every_nonzero_once():
    g, ref = _graph(700, 500, 30000, seed=5, heavy_rows=2)
    bk = _check_plan(g.fwd)
    assert (bk[:, 4] >= 0).any()                                      # the heavy rows are cut into chunks
    _check_plan(g.bwd)
    g2, _ = _graph(5000, 40, 3000, seed=6)                            # mostly empty rows: buckets of 8 rows without positions
    _check_plan(g2.fwd)
    _check_plan(g2.bwd)


@pytest.mark.parametrize("d", [64, 128, 256])
@pytest.mark.parametrize("nrhs", [1, 2])
@pytest.mark.parametrize("variant", [(0, 0, 0, 0), (0, 2, 3, 0), (0, 4, 1, 1), (0, 8, 1, 0), (0, 1, 2, 1)])
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
@pytest.mark.parametrize("tma", [0, 1])
def test_spmm_bulk_epilogues(d, tma):
    from mmssl_b200 import ops
    impl = _bulk(tma=tma)
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
    # row 0 is a heavy row (3000 non-zeros, vector reductions in arrival order): its fp32 sum moves by ~1e-6 of its magnitude
    # (~50) from run to run, which the softmax turns into ~1e-5 -- hence 3e-5 here, 5e-6 where no exponential follows
    assert rel_err(y, want) < 3e-5
    assert rel_err(s, sb.double().cpu() + want) < 3e-5
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
    _check_plan(g.fwd)
    for tma in (0, 1):
        s.copy_(s0)
        y = ops.spmm(g.fwd, [x], cs=[cc], alpha=-1.5, ss=[s], s_mode=1, impl=_bulk(tma=tma))[0]
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
            # heavy rows (~6.7k non-zeros, |logit| ~ 80) summed with float atomics in arrival order: the fp32 sum moves by ~3e-4
            # between orders and a softmax output by up to a quarter of that (7e-5 under compute-sanitizer's timing)
            assert rel_err(y, want) < 2e-4
        yt = ops.spmm(g.bwd, [torch.ones(n_rows, d, device="cuda")], impl=_bulk())[0]
        assert rel_err(yt, torch.from_numpy(np.asarray(ref.T.sum(1))).expand(-1, d)) < 1e-5
