"""One SpMM launch at the 1M x 200k / 20M-edge scale for ncu (--set full)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from mmssl_b200 import ops
from mmssl_b200.graph import BipartiteGraph
from mmssl_b200.synthetic import make_bipartite, csr_norm
U, I, nnz, d = 1_000_000, 200_000, 20_000_000, 128
r = make_bipartite(U, I, nnz, seed=1)
g = BipartiteGraph.from_scipy(csr_norm(r))
x = torch.randn(I, d, device="cuda"); y = torch.empty(U, d, device="cuda")
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for _ in range(3):
    ops.spmm(g.fwd, [x], [y], impl=impl)
torch.cuda.synchronize()
