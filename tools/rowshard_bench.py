"""Row-sharded GCN propagation on N GPUs: K-layer chain forward + backward with one
all-gather per half-layer, CUDA SpMM per row block.  Launch with torchrun; rank 0 prints one JSON line.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/rowshard_bench.py [config] [check]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from mmssl_b200 import parallel as par
from mmssl_b200.synthetic import CONFIGS, make_dataset

name = sys.argv[1] if len(sys.argv) > 1 else "syn1m"
check = "check" in sys.argv
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
U, I, nnz, d, K, _, _ = CONFIGS[name]
ds = make_dataset(name)           # every rank builds the same graph (seeded)
pu, pi = par.RowPartition(U, world), par.RowPartition(I, world)
ops_ = par.cuda_operands_from_scipy(ds.ui_norm, ds.iu_norm, pu, pi, rank, dev)
g = torch.Generator().manual_seed(0)
u0, i0 = torch.randn(U, d, generator=g) * 0.1, torch.randn(I, d, generator=g) * 0.1
gu, gi = torch.randn(U, d, generator=g), torch.randn(I, d, generator=g)
fused = "fused" in sys.argv and world > 1
if fused:
    gcn = par.FusedRowShardedGCN(ops_, pu, pi, K, rank, d, dev)
else:
    gcn = par.RowShardedGCN(ops_, pu, pi, K, par.cuda_spmm_fn, par.cuda_softmax_bwd_fn, rank)
loc = lambda t, p: p.local(t, rank).to(dev)
u0l, i0l, gul, gil = loc(u0, pu), loc(i0, pi), loc(gu, pu), loc(gi, pi)


def chain():
    s_u, s_i, saved = gcn.forward(u0l, i0l)
    return s_u, s_i, gcn.backward(saved, gul, gil)


for _ in range(3):
    out = chain()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
gcn.n_gathers = gcn.gathered_bytes = 0
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 10
a.record()
for _ in range(steps):
    out = chain()
b.record()
torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b) / steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
res = {"config": name, "n_gpus": world, "exchange": ("fused-in-SpMM " + ("multicast" if gcn.tab["u"].multicast else "peer stores")) if fused else "nccl all-gather", "ms_per_chain_fwd_bwd": round(float(ms), 4), "gathers_per_chain": gcn.n_gathers // steps,
       "gathered_MB_per_rank_per_chain": round(gcn.gathered_bytes / steps / 1e6, 2), "spmm_per_chain": 4 * K}
if check:   # parity of the sharded chain against the single-GPU engine kernels on rank 0's full graph
    s_u, s_i, (g_u0, g_i0) = out
    full = [par.all_gather_rows(t, p) for t, p in ((s_u, pu), (s_i, pi), (g_u0, pu), (g_i0, pi))]
    if rank == 0:
        one = par.RowShardedGCN(par.cuda_operands_from_scipy(ds.ui_norm, ds.iu_norm, par.RowPartition(U, 1), par.RowPartition(I, 1), 0, dev),
                                par.RowPartition(U, 1), par.RowPartition(I, 1), K, par.cuda_spmm_fn, par.cuda_softmax_bwd_fn, 0)
        su1, si1, sv = one.forward(u0.to(dev), i0.to(dev))
        gu1, gi1 = one.backward(sv, gu.to(dev), gi.to(dev))
        errs = [float((x - y).abs().max() / y.abs().max()) for x, y in zip(full, (su1, si1, gu1, gi1))]
        res["max_rel_err_vs_1gpu"] = max(errs)
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
