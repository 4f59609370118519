import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, scipy.sparse as sp, torch, ctypes as C
from mmssl_b200 import ops, _lib
from mmssl_b200.graph import BipartiteGraph
from mmssl_b200._lib import ptr, stream, SpmmRhs
rng = np.random.default_rng(1)
n_rows, n_cols, nnz, d = 3000, 5000, 90000, 64
pw = 1.0 / np.arange(1, n_cols + 1); pw /= pw.sum()
r = rng.integers(0, n_rows, nnz); c = rng.choice(n_cols, nnz, p=pw); v = rng.standard_normal(nnz).astype(np.float32)
ref = sp.coo_matrix((v.astype(np.float64), (r, c)), shape=(n_rows, n_cols)).tocsr()
g = BipartiteGraph(torch.from_numpy(r).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(v).cuda(), (n_rows, n_cols))
x = torch.randn(n_cols, d, device="cuda")
want = torch.from_numpy(ref @ x.double().cpu().numpy())
lib = _lib.load(True)
a = g.fwd
colidx_hot, hot_ids, n_hot = a.hot_plan()
for nh in (0, 1, 8, 100, 500, 799, 2048):
    y = torch.empty(n_rows, d, device="cuda")
    rhs = (SpmmRhs * 1)(SpmmRhs(ptr(x), x.stride(0), ptr(y), y.stride(0), None, 0, None, 0, None, 0, None, 0))
    part, cnt = a.work_area(d)
    desc = type(a.desc).from_buffer_copy(a.desc); desc.counters = cnt.data_ptr()
    # bypass the d>=128 guard by calling with the internal symbol? not exported: temporarily allowed via env
    rc = lib.mmssl_spmm_hot_f32(C.byref(desc), ptr(colidx_hot), ptr(hot_ids), min(nh, n_hot), d, 1, rhs, 0, 1.0, 0, ptr(part), part.numel(), stream())
    if rc:
        print("rc", rc, lib.mmssl_last_error()); break
    torch.cuda.synchronize()
    err = (y.double().cpu() - want).abs()
    bad_rows = (err.max(1).values > 1e-4).nonzero().flatten()
    print("n_hot", nh, "max err", float(err.max()), "bad rows", len(bad_rows), bad_rows[:8].tolist())
