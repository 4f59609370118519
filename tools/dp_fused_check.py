"""Fused DP optimiser (multimem reduce-scatter + sharded AdamW + multimem all-gather, one kernel) vs the
NCCL all-reduce + replicated AdamW path: parity and time.  torchrun --nproc-per-node N tools/dp_fused_check.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from mmssl_b200 import ops, parallel as par
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shapes = {"wv": (64, 4096), "bv": (64,), "wt": (64, 1024), "bt": (64,), "eu": (19445, 64), "ei": (7050, 64), "wc": (256, 64)}
g0 = torch.Generator().manual_seed(0)
P = {k: torch.randn(*s, generator=g0).to(dev) for k, s in shapes.items()}
# reference path
ref = {k: v.clone() for k, v in P.items()}
bucket = par.GradBucket({k: torch.zeros_like(v) for k, v in ref.items()})
m = {k: torch.zeros_like(v) for k, v in ref.items()}; vv = {k: torch.zeros_like(v) for k, v in ref.items()}
step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
opt = par.FusedDPOptimizer(P, rank, world, lr=5.5e-4)
keys = list(shapes)
def grads_for(step):
    g = torch.Generator().manual_seed(1000 * step + rank)
    return {k: torch.randn(*s, generator=g).to(dev) for k, s in shapes.items()}
for step in range(1, 4):
    gs = grads_for(step)
    for k in keys:
        bucket.views[k].copy_(gs[k]); opt.grads[k].copy_(gs[k])
    bucket.all_reduce_mean()
    ops.step_tick(step_dev)
    ops.adamw([ref[k] for k in keys], [bucket.views[k] for k in keys], [m[k] for k in keys], [vv[k] for k in keys], step_dev, 5.5e-4)
    opt.step()
torch.cuda.synchronize()
err = max(float((opt.params[k] - ref[k]).abs().max() / ref[k].abs().max()) for k in keys)
def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps * 1e3], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t)
def nccl_path():
    bucket.all_reduce_mean(); ops.step_tick(step_dev)
    ops.adamw([ref[k] for k in keys], [bucket.views[k] for k in keys], [m[k] for k in keys], [vv[k] for k in keys], step_dev, 5.5e-4)
t_nccl = timeit(nccl_path); t_fused = timeit(opt.step)
if rank == 0:
    print(json.dumps({"n_gpus": world, "params": sum(v.numel() for v in P.values()), "max_rel_err_vs_nccl_path": err,
                      "nccl_allreduce_plus_adamw_us": round(t_nccl, 1), "fused_multimem_us": round(t_fused, 1)}))
dist.destroy_process_group()
