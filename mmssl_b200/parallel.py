"""Multi-GPU plumbing for the MMSSL hot path.  One process per GPU,
``torch.distributed`` (NCCL over NVLink on the box, gloo in the CPU tests) for the exchanges.

Two schemes:

* **Data parallel over the batch** (`GradBucket`): graph, tables and features replicated; each rank
  runs the hot step on its own triples; the gradients of the live parameters live in ONE flat
  buffer that is all-reduced (mean) once per step; AdamW is replicated.  Used by ``bench.py --gpus N``.

* **Row-sharded propagation** (`RowShardedGCN`): user rows and item rows are block-partitioned over
  the ranks.  ``u = A_ui i`` needs the full ``i`` -> one all-gather of the ``[rows/G, d]`` slice the
  previous SpMM produced; because items use the *new* users inside a layer (Models.py:203-208) that
  is two all-gathers per GCN layer, and the same count in backward where ``dX = A^T dY`` is computed
  row-locally from the gathered ``dY`` (deterministic: no reduce-scatter, no float atomics across
  ranks).  The SpMM itself is injected (`spmm_fn`): the CUDA operator in production, a CPU
  restatement in the gloo tests -- this module only owns partitioning and the exchange schedule.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

EPI_NONE, EPI_SOFTMAX, EPI_SOFTMAX_BWD = 0, 1, 2


# ------------------------------------------------------------------------------------------ DP
class GradBucket:
    """Re-points a dict of gradient tensors into one flat buffer (16-byte aligned slices) so that a
    single all-reduce covers all live parameters."""

    def __init__(self, grads: Dict[str, torch.Tensor]):
        self.keys = list(grads.keys())
        offs, total = [], 0
        for k in self.keys:
            offs.append(total)
            total += (grads[k].numel() + 3) // 4 * 4
        like = grads[self.keys[0]]
        self.flat = torch.zeros(total, dtype=like.dtype, device=like.device)
        self.views: Dict[str, torch.Tensor] = {}
        for k, o in zip(self.keys, offs):
            self.views[k] = self.flat[o:o + grads[k].numel()].view_as(grads[k])
            self.views[k].copy_(grads[k])

    def all_reduce_mean(self, group=None) -> None:
        world = dist.get_world_size(group)
        if world == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.mul_(1.0 / world)


class FusedDPOptimizer:
    """Data-parallel optimiser step as ONE kernel over NVSwitch multicast (mmssl_dp_fused_adamw):
    reduce-scatter of the gradient bucket by `multimem.ld_reduce` (summed in the switch), AdamW on this
    rank's 1/world slice (m, v are sharded), all-gather of the new parameters by `multimem.st`.
    Parameters and gradients live in two symmetric-memory buckets with identical layout on every rank;
    `params` / `grads` are views into them.  No NCCL on the data path; two signal-pad barriers per step."""

    def __init__(self, params: Dict[str, torch.Tensor], rank: int, world: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8,
                 weight_decay=1e-2, group=None):
        import torch.distributed._symmetric_memory as symm
        self.keys = list(params.keys())
        self.rank, self.world = rank, world
        self.hyper = (float(lr), float(beta1), float(beta2), float(eps), float(weight_decay))
        offs, total = [], 0
        for k in self.keys:
            offs.append(total)
            total += (params[k].numel() + 3) // 4 * 4
        unit = 4 * world
        total = (total + unit - 1) // unit * unit
        dev = params[self.keys[0]].device
        g = group if group is not None else dist.group.WORLD
        self.pflat = symm.empty(total, dtype=torch.float32, device=dev)
        self.gflat = symm.empty(total, dtype=torch.float32, device=dev)
        self.pflat.zero_(); self.gflat.zero_()
        self.hp = symm.rendezvous(self.pflat, g.group_name)
        self.hg = symm.rendezvous(self.gflat, g.group_name)
        self.p_mc, self.g_mc = int(self.hp.multicast_ptr), int(self.hg.multicast_ptr)
        if self.p_mc == 0 or self.g_mc == 0:
            raise RuntimeError("NVSwitch multicast is not available for symmetric memory on this system")
        self.params: Dict[str, torch.Tensor] = {}
        self.grads: Dict[str, torch.Tensor] = {}
        for k, o in zip(self.keys, offs):
            n = params[k].numel()
            self.params[k] = self.pflat[o:o + n].view_as(params[k])
            self.params[k].copy_(params[k])
            self.grads[k] = self.gflat[o:o + n].view_as(params[k])
        self.count = total // world
        self.begin = rank * self.count
        self.m = torch.zeros(self.count, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.count, dtype=torch.float32, device=dev)
        self.step_no = 0
        torch.cuda.synchronize(dev)
        self.hp.barrier()

    def step(self) -> None:
        from . import _lib
        from ._lib import ptr, stream
        lib = _lib.load(require_device=True)
        self.step_no += 1
        lr, b1, b2, eps, wd = self.hyper
        self.hg.barrier()          # every rank's gradients are complete
        _lib.check(lib.mmssl_dp_fused_adamw(ptr(self.pflat), self.p_mc, self.g_mc, ptr(self.m), ptr(self.v), self.begin,
                                            self.count, 1.0 / self.world, self.step_no, lr, b1, b2, eps, wd, stream()))
        self.hp.barrier()          # every slice of the new parameters has been published

    def step_captured(self, step_dev: torch.Tensor) -> None:
        """The same step with the step number read from the device counter `step_dev` (int32, already ticked): two signal-pad
        barriers + one kernel, nothing host-dependent -- meant to be captured INSIDE the step's CUDA graph (the
        eager launches after the replay were what the N >= 2 step paid over the 1-GPU step)."""
        from . import _lib
        from ._lib import ptr, stream
        lib = _lib.load(require_device=True)
        lr, b1, b2, eps, wd = self.hyper
        self.hg.barrier()
        _lib.check(lib.mmssl_dp_fused_adamw_dev(ptr(self.pflat), self.p_mc, self.g_mc, ptr(self.m), ptr(self.v), self.begin,
                                                self.count, 1.0 / self.world, ptr(step_dev), lr, b1, b2, eps, wd, stream()))
        self.hp.barrier()


# ------------------------------------------------------------------------------------------ row sharding
@dataclass(frozen=True)
class RowPartition:
    """Equal blocks of ``block = ceil(n / world)`` rows (the last one padded) so that slices can be
    exchanged with ``all_gather_into_tensor``."""
    n: int
    world: int

    @property
    def block(self) -> int:
        return (self.n + self.world - 1) // self.world

    def bounds(self, rank: int) -> Tuple[int, int]:
        lo = min(self.n, rank * self.block)
        return lo, min(self.n, lo + self.block)

    def local(self, full: torch.Tensor, rank: int) -> torch.Tensor:
        """Padded [block, d] slice of a full [n, d] table."""
        lo, hi = self.bounds(rank)
        out = torch.zeros(self.block, full.shape[1], dtype=full.dtype, device=full.device)
        out[:hi - lo] = full[lo:hi]
        return out


def all_gather_rows(local: torch.Tensor, part: RowPartition, group=None) -> torch.Tensor:
    """[block, d] per rank -> full [n, d] (padding rows dropped)."""
    if part.world == 1:
        return local[:part.n]
    out = torch.empty(part.world * part.block, local.shape[1], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:part.n]


def shard_rows_scipy(mat, part: RowPartition, rank: int):
    """Row block of a scipy sparse matrix, padded to ``part.block`` rows (global column ids kept)."""
    import scipy.sparse as sp
    lo, hi = part.bounds(rank)
    blk = mat.tocsr()[lo:hi]
    if hi - lo < part.block:
        blk = sp.vstack([blk, sp.csr_matrix((part.block - (hi - lo), mat.shape[1]), dtype=blk.dtype)]).tocsr()
    return blk


class RowShardedGCN:
    """K-layer user<->item propagation (Models.py:199-211) over row-sharded tables.

    operands: dict with the rank's row blocks  'ui' = A_ui[U_r, :], 'iu' = A_iu[I_r, :],
              'iuT' = (A_iu^T)[U_r, :], 'uiT' = (A_ui^T)[I_r, :]   (whatever type `spmm_fn` accepts)
    spmm_fn(op, x_full, *, c=None, alpha=1.0, epilogue=EPI_NONE, ysaved=None) -> y_local [block, d]
    softmax_bwd_fn(y, g, alpha) -> y * (alpha*g - <alpha*g, y>)
    """

    def __init__(self, operands: Dict[str, object], part_u: RowPartition, part_i: RowPartition, n_layers: int,
                 spmm_fn: Callable, softmax_bwd_fn: Callable, rank: int, group=None):
        self.ops, self.pu, self.pi, self.K = operands, part_u, part_i, n_layers
        self.spmm, self.softmax_bwd = spmm_fn, softmax_bwd_fn
        self.rank, self.group = rank, group
        self.n_gathers = 0
        self.gathered_bytes = 0

    def _gather(self, local, part):
        self.n_gathers += 1
        self.gathered_bytes += local.numel() * local.element_size() * (part.world - 1)
        return all_gather_rows(local, part, self.group)

    def forward(self, u0_local: torch.Tensor, i0_local: torch.Tensor):
        """Returns (S_u_local, S_i_local, saved): S = u_0 + ... + u_K on the rank's rows."""
        s_u, s_i = u0_local.clone(), i0_local.clone()
        cur_i = self._gather(i0_local, self.pi)
        u_last = i_last = None
        for k in range(self.K):
            last = k == self.K - 1
            epi = EPI_SOFTMAX if last else EPI_NONE
            u_loc = self.spmm(self.ops["ui"], cur_i, epilogue=epi)
            s_u += u_loc
            u_full = self._gather(u_loc, self.pu)
            i_loc = self.spmm(self.ops["iu"], u_full, epilogue=epi)
            s_i += i_loc
            if last:
                u_last, i_last = u_loc, i_loc
            else:
                cur_i = self._gather(i_loc, self.pi)
        return s_u, s_i, (u_last, i_last)

    def backward(self, saved, g_su_local: torch.Tensor, g_si_local: torch.Tensor):
        """g_s*_local: gradient w.r.t. every u_k / i_k on the rank's rows (= g_uf / (K+1), g_if / (K+1)).
        Returns (g_u0_local, g_i0_local)."""
        u_last, i_last = saved
        if self.K == 0:
            return g_su_local, g_si_local
        t_loc = self.softmax_bwd(i_last, g_si_local, 1.0)
        for k in range(self.K - 1, -1, -1):
            last = k == self.K - 1
            t_full = self._gather(t_loc, self.pi)
            tu_loc = self.spmm(self.ops["iuT"], t_full, c=g_su_local, alpha=1.0,
                               epilogue=EPI_SOFTMAX_BWD if last else EPI_NONE, ysaved=u_last if last else None)
            tu_full = self._gather(tu_loc, self.pu)
            t_loc = self.spmm(self.ops["uiT"], tu_full, c=g_si_local, alpha=1.0)
        return g_su_local, t_loc


# ------------------------------------------------------------------------------------------ fused SpMM + all-gather
class SymmetricTable:
    """A full [world*block, d] fp32 table that exists at the same place on every rank (CUDA symmetric
    memory).  A rank's SpMM writes its row block straight into EVERY rank's copy from the kernel
    epilogue -- one `multimem.st` per 16 bytes through the NVSwitch multicast address when available,
    else one NVLink store per peer -- so the all-gather is part of the SpMM and no NCCL call is made;
    `barrier()` (device-side signal pads) orders producers and consumers."""

    def __init__(self, part: RowPartition, d: int, rank: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.part, self.d, self.rank = part, d, rank
        self.t = symm.empty(part.world * part.block, d, dtype=torch.float32, device=device)
        self.t.zero_()
        g = group if group is not None else dist.group.WORLD
        self.h = symm.rendezvous(self.t, g.group_name)
        self.ptrs = [int(p) for p in self.h.buffer_ptrs]
        mc = 0
        try:
            mc = int(self.h.multicast_ptr)
        except Exception:
            mc = 0
        self.mc = mc
        self.row_bytes = d * 4

    @property
    def multicast(self) -> bool:
        return self.mc != 0

    def local_rows(self) -> torch.Tensor:
        lo = self.rank * self.part.block
        return self.t[lo:lo + self.part.block]

    def full(self) -> torch.Tensor:
        return self.t[:self.part.n]

    def out_spec(self):
        """kwargs for ops.spmm so that the output rows land in every rank's table."""
        off = self.rank * self.part.block * self.row_bytes
        if self.multicast:
            return dict(y_mode=1, y_raw=[self.mc + off])
        return dict(y_mode=2, y_peers=[[p + off for r, p in enumerate(self.ptrs) if r != self.rank]])

    def barrier(self):
        self.h.barrier()


class FusedRowShardedGCN(RowShardedGCN):
    """RowShardedGCN whose exchanges are fused into the SpMM epilogue (no NCCL on the data path)."""

    def __init__(self, operands, part_u, part_i, n_layers, rank, d, device, group=None):
        super().__init__(operands, part_u, part_i, n_layers, cuda_spmm_fn, cuda_softmax_bwd_fn, rank, group)
        mk = lambda part: SymmetricTable(part, d, rank, device, group)
        self.tab = {"u": mk(part_u), "i": mk(part_i), "gu": mk(part_u), "gi": mk(part_i)}

    def _spmm_into(self, op, x_full, table: SymmetricTable, **kw):
        from . import ops
        y_local = table.local_rows()
        ops.spmm(op, [x_full.contiguous()], [y_local], cs=[kw["c"]] if kw.get("c") is not None else None,
                 alpha=kw.get("alpha", 1.0), epilogue=kw.get("epilogue", EPI_NONE),
                 ysaved=[kw["ysaved"]] if kw.get("ysaved") is not None else None, **table.out_spec())
        table.barrier()
        self.n_gathers += 1
        self.gathered_bytes += y_local.numel() * 4 * (table.part.world - 1)
        return y_local, table.full()

    def forward(self, u0_local, i0_local):
        s_u, s_i = u0_local.clone(), i0_local.clone()
        cur_i = self._gather(i0_local, self.pi)            # i_0 is not an SpMM output: plain all-gather
        u_last = i_last = None
        for k in range(self.K):
            last = k == self.K - 1
            epi = EPI_SOFTMAX if last else EPI_NONE
            u_loc, u_full = self._spmm_into(self.ops["ui"], cur_i, self.tab["u"], epilogue=epi)
            s_u += u_loc
            i_loc, cur_i = self._spmm_into(self.ops["iu"], u_full, self.tab["i"], epilogue=epi)
            s_i += i_loc
            if last:
                u_last, i_last = u_loc.clone(), i_loc.clone()
        return s_u, s_i, (u_last, i_last)

    def backward(self, saved, g_su_local, g_si_local):
        u_last, i_last = saved
        if self.K == 0:
            return g_su_local, g_si_local
        t_loc = self.softmax_bwd(i_last, g_si_local, 1.0)
        t_full = self._gather(t_loc, self.pi)               # first exchange: not an SpMM output
        for k in range(self.K - 1, -1, -1):
            last = k == self.K - 1
            _, tu_full = self._spmm_into(self.ops["iuT"], t_full, self.tab["gu"], c=g_su_local, alpha=1.0,
                                         epilogue=EPI_SOFTMAX_BWD if last else EPI_NONE, ysaved=u_last if last else None)
            t_loc, t_full = self._spmm_into(self.ops["uiT"], tu_full, self.tab["gi"], c=g_si_local, alpha=1.0)
        return g_su_local, t_loc.clone()


# ------------------------------------------------------------------------------------------ CUDA binding
def cuda_operands_from_scipy(ui_norm, iu_norm, part_u: RowPartition, part_i: RowPartition, rank: int, device):
    """Row blocks of A_ui, A_iu and of their transposes as prepared CUDA SpMM operands."""
    from .graph import SparseOperand

    def op(mat, part):
        blk = shard_rows_scipy(mat, part, rank).tocoo()
        r = torch.from_numpy(blk.row.astype("int64")).to(device)
        c = torch.from_numpy(blk.col.astype("int64")).to(device)
        v = torch.from_numpy(blk.data.astype("float32")).to(device)
        o = SparseOperand(r, c, v, blk.shape[0], blk.shape[1])
        o.tighten()
        return o

    return {"ui": op(ui_norm, part_u), "iu": op(iu_norm, part_i),
            "iuT": op(iu_norm.T.tocsr(), part_u), "uiT": op(ui_norm.T.tocsr(), part_i)}


def cuda_spmm_fn(op, x_full, *, c=None, alpha=1.0, epilogue=EPI_NONE, ysaved=None):
    from . import ops
    return ops.spmm(op, [x_full.contiguous()], cs=[c] if c is not None else None, alpha=alpha, epilogue=epilogue,
                    ysaved=[ysaved] if ysaved is not None else None)[0]


def cuda_softmax_bwd_fn(y, g, alpha):
    from . import ops
    return ops.softmax_bwd(y, g, alpha, torch.empty_like(y))
