"""Prepared sparse operands for the SpMM operator.

The reference hands ``MMSSL.forward`` six torch sparse COO tensors (int64 indices, fp32 values,
flagged uncoalesced; built by ``matrix_to_tensor`` / ``sparse_mx_to_torch_sparse_tensor``,
main.py:105-112, :513-520) and lets ATen re-coalesce and convert them on every ``torch.sparse.mm``.
Here each distinct tensor is converted ONCE, on the device, into
  * CSR of A     (forward products  A @ X),
  * CSR of A^T   (backward products A^T @ dY; A_ui^T != A_iu because each side is normalised by
                  its own row degree, main.py:66-67),
  * an nnz-balanced work plan for each,
and cached by tensor identity (the static ui/iu graphs live for the whole run; the modality graphs
are rebuilt by the trainer and simply miss the cache).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import CsrDesc, ptr, stream


_cuts_applied = False


def _env_cuts(lib) -> None:
    """MMSSL_SPMM_CUTS="split_threshold,seg_len,heavy_threshold,heavy_seg_len": experiment knob for where the plan cuts rows."""
    global _cuts_applied
    if not _cuts_applied:
        _cuts_applied = True
        v = os.environ.get("MMSSL_SPMM_CUTS")
        if v:
            _lib.check(lib.mmssl_spmm_plan_set_cuts(*[int(x) for x in v.split(",")]))


class SparseOperand:
    """CSR matrix + SpMM work plan living on one CUDA device."""

    def __init__(self, rows: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, n_rows: int, n_cols: int,
                 transpose: bool = False):
        lib = _lib.load(require_device=True)
        dev = vals.device
        nnz = int(vals.numel())
        if transpose:
            n_rows, n_cols = n_cols, n_rows
        self.n_rows, self.n_cols, self.nnz = int(n_rows), int(n_cols), nnz
        self.device = dev
        i32 = dict(dtype=torch.int32, device=dev)
        self.rowptr = torch.empty(n_rows + 1, **i32)
        self.colidx = torch.empty(max(nnz, 1), **i32)
        self.vals = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
        ws_bytes = lib.mmssl_csr_workspace_bytes(nnz, n_rows)
        if ws_bytes < 0:
            raise _lib.MmsslLibraryError("mmssl_csr_workspace_bytes failed")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        rows = rows.contiguous()
        cols = cols.contiguous()
        vals = vals.contiguous().to(torch.float32)
        _lib.check(lib.mmssl_csr_from_coo(ptr(rows), ptr(cols), ptr(vals), nnz, n_rows, n_cols, int(transpose),
                                          ptr(self.rowptr), ptr(self.colidx), ptr(self.vals), ptr(ws), ws_bytes, stream()))
        # work plan
        _env_cuts(lib)
        self.items_cap = lib.mmssl_spmm_plan_items_cap(n_rows, nnz)
        self.splits_cap = lib.mmssl_spmm_plan_splits_cap(nnz)
        self.segs_cap = lib.mmssl_spmm_plan_segs_cap(nnz)
        self.items = torch.empty(self.items_cap * 4, **i32)
        self.split_table = torch.empty(self.splits_cap * 4, **i32)
        self.counters = torch.empty(self.splits_cap, **i32)
        self.totals = torch.empty(3, **i32)
        pws_bytes = lib.mmssl_spmm_plan_workspace_bytes(n_rows)
        pws = torch.empty(pws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.mmssl_spmm_plan(ptr(self.rowptr), n_rows, nnz, ptr(self.items), self.items_cap,
                                       ptr(self.split_table), ptr(self.counters), self.splits_cap, ptr(self.totals),
                                       ptr(pws), pws_bytes, stream()))
        self._keepalive = (ws, pws, rows, cols, vals)   # until the stream has consumed them
        # without split rows the launch only needs n_rows items; with them we launch over the
        # capacity (unused entries are row=-1).  Resolved lazily (first host read of `totals`).
        self._n_items_exact: Optional[int] = None
        self._work = {}
        self._hot = None
        self._bulk = None
        self.desc = CsrDesc(ptr(self.rowptr), ptr(self.colidx), ptr(self.vals), n_rows, n_cols, nnz, ptr(self.items),
                            self.items_cap, ptr(self.split_table), ptr(self.counters), self.segs_cap)
        if os.environ.get("MMSSL_SPMM_SORT") == "1" and nnz > 0:
            self.sort_items_by_length()

    def work_area(self, width: int):
        """(partials, counters) for launches whose right-hand sides total `width` floats per row.
        Zero-initialised once; the kernel leaves counters and heavy-row slots clean after every launch."""
        w = self._work.get(width)
        if w is None:
            part = torch.zeros(max(self.segs_cap * width, 4), dtype=torch.float32, device=self.device)
            cnt = torch.zeros(max(self.splits_cap, 1), dtype=torch.int32, device=self.device)
            w = (part, cnt)
            self._work[width] = w
        return w

    def bulk_plan(self):
        """Bucket plan of the staged-gather SpMM (csrc/spmm_bulk.cu): buckets of <= 32 non-zeros / <= 8 whole rows, rows over 32
        non-zeros cut into 32-chunks.  Built once per operand on first use (one host read of the bucket count)."""
        if self._bulk is None:
            lib = _lib.load(require_device=True)
            i32 = dict(dtype=torch.int32, device=self.device)
            splits_cap = lib.mmssl_spmm_bulk_plan_splits_cap(self.nnz)
            segs_cap = lib.mmssl_spmm_bulk_plan_segs_cap(self.nnz)
            buckets_cap = lib.mmssl_spmm_bulk_plan_buckets_cap(self.n_rows, self.nnz)
            split_table = torch.empty(splits_cap * 4, **i32)
            counters = torch.empty(splits_cap, **i32)
            buckets = torch.empty(buckets_cap * 8, **i32)
            totals = torch.empty(3, **i32)
            pws_bytes = lib.mmssl_spmm_bulk_plan_workspace_bytes(self.n_rows)
            pws = torch.empty(pws_bytes, dtype=torch.uint8, device=self.device)
            _lib.check(lib.mmssl_spmm_bulk_plan(ptr(self.rowptr), self.n_rows, self.nnz, ptr(split_table), ptr(counters), splits_cap,
                                                ptr(buckets), buckets_cap, ptr(totals), ptr(pws), pws_bytes, stream()))
            n_buckets, n_split, n_segs = (int(x) for x in totals.cpu())
            buckets = buckets[:max(n_buckets, 1) * 8].clone()             # exact size (the capacity bound is generous)
            desc = CsrDesc(ptr(self.rowptr), ptr(self.colidx), ptr(self.vals), self.n_rows, self.n_cols, self.nnz, ptr(self.rowptr),
                           0, ptr(split_table), ptr(counters), max(n_segs, 0))
            self._bulk = dict(desc=desc, buckets=buckets, n_buckets=n_buckets, split_table=split_table, counters=counters,
                              totals=totals, segs_cap=max(n_segs, 0), splits_cap=splits_cap, n_split=n_split, work={})
        return self._bulk

    def bulk_work_area(self, width: int):
        """(partials, counters) of the bulk plan for launches whose right-hand sides total `width` floats per row."""
        b = self.bulk_plan()
        w = b["work"].get(width)
        if w is None:
            part = torch.zeros(max(b["segs_cap"] * width, 4), dtype=torch.float32, device=self.device)
            cnt = torch.zeros(max(b["splits_cap"], 1), dtype=torch.int32, device=self.device)
            w = (part, cnt)
            b["work"][width] = w
        return w

    def hot_flag_colidx(self, n_hot: int):
        """Copy of the column index array with the sign bit set on the `n_hot` highest-degree columns (LDG kernel, impl bit 8:
        hot rows are pinned in L1, cold rows bypass it).  One-time graph preparation, device side."""
        key = ("hotflag", n_hot)
        if key not in self._work:
            col = self.colidx[:self.nnz].long()
            deg = torch.bincount(col, minlength=self.n_cols)
            h = int(min(n_hot, self.n_cols))
            flag = torch.zeros(self.n_cols, dtype=torch.bool, device=self.device)
            if h > 0 and self.nnz > 0:
                flag[torch.topk(deg, h).indices] = True
            buf = self.colidx.clone()
            if self.nnz > 0:
                buf[:self.nnz] = torch.where(flag[col], self.colidx[:self.nnz] | -2147483648, self.colidx[:self.nnz])
                self.hot_flag_fraction = float(flag[col].float().mean())
            self._work[key] = buf
        return self._work[key]

    def hot_plan(self, max_slots: int = 2048):
        """(colidx_hot, hot_ids, n_hot) for the TMA-staged SpMM: the `max_slots` highest-degree columns
        get shared-memory slots (ordered by decreasing degree) and are encoded as -(slot+1) in a copy of
        the column index array.  One-time graph preparation (device-side index manipulation)."""
        if self._hot is None:
            col = self.colidx[:self.nnz].long()
            deg = torch.bincount(col, minlength=self.n_cols)
            h = int(min(max_slots, self.n_cols, int((deg > 0).sum())))
            hot_ids = torch.topk(deg, h).indices.to(torch.int32) if h > 0 else torch.zeros(1, dtype=torch.int32, device=self.device)
            slot = torch.full((self.n_cols,), -1, dtype=torch.int32, device=self.device)
            if h > 0:
                slot[hot_ids.long()] = torch.arange(h, dtype=torch.int32, device=self.device)
            sl = slot[col]
            hot_col = torch.where(sl >= 0, -(sl + 1), self.colidx[:self.nnz])
            buf = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=self.device)
            buf[:self.nnz] = hot_col
            self._hot = (buf, hot_ids.contiguous(), h)
            self.hot_edge_fraction = float((sl >= 0).float().mean()) if self.nnz else 0.0
        return self._hot

    def tighten(self) -> None:
        """Optional: read the exact item count back (one host sync) so launches are not padded."""
        if self._n_items_exact is None:
            t = self.totals.cpu()
            self._n_items_exact = int(t[0])
            self.n_split_rows, self.n_segs = int(t[1]), int(t[2])
            self.desc.n_items = self._n_items_exact
            self.desc.segs_cap = max(self.n_segs, 0)
            self.segs_cap = self.n_segs
            self._keepalive = ()

    def sort_items_by_length(self) -> None:
        """Candidate awaiting measurement (MMSSL_SPMM_SORT=1): reorder the work items longest first.  Every item is
        self-contained (row, [begin, end), split slot), so any order gives the same result; what changes is who shares a warp
        (at d = 64 two items do: similar lengths waste fewer lanes) and that the long items start first (shorter tail)."""
        self.tighten()
        n = self._n_items_exact
        if n <= 1:
            return
        it = self.items[:4 * n].view(n, 4)
        order = torch.argsort(it[:, 2] - it[:, 1], descending=True, stable=True)
        self.items[:4 * n] = it[order].reshape(-1)

    def to_scipy(self):
        import numpy as np
        import scipy.sparse as sp
        return sp.csr_matrix((self.vals[:self.nnz].cpu().numpy(), self.colidx[:self.nnz].cpu().numpy(),
                              self.rowptr.cpu().numpy().astype(np.int64)), shape=(self.n_rows, self.n_cols))


class BipartiteGraph:
    """A and A^T prepared for SpMM.  ``fwd`` multiplies by A, ``bwd`` by A^T."""

    def __init__(self, rows, cols, vals, shape: Tuple[int, int], tighten: bool = True):
        self.shape = (int(shape[0]), int(shape[1]))
        self.nnz = int(vals.numel())
        self.fwd = SparseOperand(rows, cols, vals, shape[0], shape[1], transpose=False)
        self.bwd = SparseOperand(rows, cols, vals, shape[0], shape[1], transpose=True)
        if tighten:
            self.fwd.tighten()
            self.bwd.tighten()

    @classmethod
    def from_torch_sparse(cls, t: torch.Tensor, tighten: bool = True) -> "BipartiteGraph":
        if t.layout != torch.sparse_coo:
            raise TypeError("expected a torch sparse COO tensor (as built by the reference's matrix_to_tensor)")
        if not t.is_cuda:
            raise _lib.MmsslLibraryError("graph tensors must live on the CUDA device; there is no CPU path")
        idx = t._indices()
        return cls(idx[0], idx[1], t._values(), tuple(t.shape), tighten=tighten)

    @classmethod
    def from_scipy(cls, mat, device="cuda", tighten: bool = True) -> "BipartiteGraph":
        coo = mat.tocoo()
        rows = torch.from_numpy(coo.row.astype("int64")).to(device)
        cols = torch.from_numpy(coo.col.astype("int64")).to(device)
        vals = torch.from_numpy(coo.data.astype("float32")).to(device)
        return cls(rows, cols, vals, coo.shape, tighten=tighten)


# ---- cache keyed by tensor identity -------------------------------------------------------------
_cache: Dict[int, Tuple[weakref.ref, BipartiteGraph]] = {}


def prepare(t) -> BipartiteGraph:
    """BipartiteGraph for a torch sparse tensor (cached by object identity) or pass-through."""
    if isinstance(t, BipartiteGraph):
        return t
    key = id(t)
    hit = _cache.get(key)
    if hit is not None and hit[0]() is t:
        return hit[1]
    g = BipartiteGraph.from_torch_sparse(t)
    if len(_cache) > 64:   # drop entries whose tensors died
        for k in [k for k, (r, _) in _cache.items() if r() is None]:
            del _cache[k]
    _cache[key] = (weakref.ref(t), g)
    return g
