"""Synthetic datasets with the shapes BASELINE.json names (no network: the real Amazon / Tiktok
files are on Google Drive, README.md:52 of the reference).

Graph: bipartite R in {0,1}^{U x I} with exactly `nnz` edges; user degrees ~ lognormal(sigma=1)
(>= 1, rescaled to nnz), item endpoints ~ Zipf(alpha) over a random permutation of the items
(popularity skew), duplicate (u, i) pairs re-drawn.  Normalisation = the reference's
``csr_norm(mean_flag=True)`` (main.py:89-103): D_row^{-1/2} A for A = R and, separately, A = R^T.
Sampler: the semantics of ``Data.sample`` (utility/load_data.py:153-191), vectorised.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np
import scipy.sparse as sp

CONFIGS = {
    # name: (U, I, nnz, d, n_layers, Dv, Dt)   -- BASELINE.json configs[0..4], README.md:40-47 shapes
    "tiktok": (9319, 6710, 59541, 64, 2, 128, 768),
    "baby": (19445, 7050, 139110, 64, 2, 4096, 1024),
    "sports": (35598, 18357, 256308, 64, 3, 4096, 1024),
    "syn1m": (1_000_000, 200_000, 20_000_000, 128, 2, 4096, 1024),
    "syn10m": (10_000_000, 1_000_000, 200_000_000, 256, 2, 4096, 1024),
    "tiny": (300, 200, 2500, 64, 2, 96, 40),
}


def make_bipartite(n_users: int, n_items: int, nnz: int, seed: int = 2022, alpha: float = 1.0,
                   uniform: bool = False) -> sp.csr_matrix:
    rng = np.random.default_rng(seed)
    if uniform:
        deg = np.full(n_users, nnz // n_users, np.int64)
    else:
        w = rng.lognormal(0.0, 1.0, n_users)
        deg = np.maximum(1, np.floor(w / w.sum() * nnz)).astype(np.int64)
    deg = np.minimum(deg, max(1, n_items // 4))
    # fix the total to exactly nnz
    diff = int(nnz - deg.sum())
    while diff != 0:
        step = min(abs(diff), n_users)
        pick = rng.choice(n_users, size=step, replace=False)
        if diff > 0:
            ok = pick[deg[pick] < max(1, n_items // 4)]
            deg[ok] += 1
            diff -= len(ok)
        else:
            ok = pick[deg[pick] > 1]
            deg[ok] -= 1
            diff += len(ok)
    perm = rng.permutation(n_items)
    if uniform:
        cdf = np.arange(1, n_items + 1, dtype=np.float64) / n_items
    else:
        pw = 1.0 / np.power(np.arange(1, n_items + 1, dtype=np.float64), alpha)
        cdf = np.cumsum(pw / pw.sum())
    users = np.repeat(np.arange(n_users, dtype=np.int64), deg)

    def draw(k):
        return perm[np.minimum(np.searchsorted(cdf, rng.random(k)), n_items - 1)]

    items = draw(len(users))
    keys = users * n_items + items
    for it in range(200):
        _, first = np.unique(keys, return_index=True)
        dup = np.ones(len(keys), bool)
        dup[first] = False
        n_dup = int(dup.sum())
        if n_dup == 0:
            break
        # popular items saturate quickly: flatten the redraw distribution progressively
        items[dup] = draw(n_dup) if it < 20 else rng.integers(0, n_items, n_dup)
        keys = users * n_items + items
    else:
        raise RuntimeError("could not place all edges without duplicates")
    mat = sp.csr_matrix((np.ones(len(users), np.float32), (users, items)), shape=(n_users, n_items))
    assert mat.nnz == nnz, (mat.nnz, nnz)
    return mat


def csr_norm(mat: sp.spmatrix) -> sp.csr_matrix:
    """csr_norm(mean_flag=True) of the reference (main.py:89-103): D_row^{-1/2} A, +1e-8 inside the power."""
    rs = np.asarray(mat.sum(1)).ravel()
    scale = np.power(rs + 1e-8, -0.5)
    scale[np.isinf(scale)] = 0.0
    return (sp.diags(scale) * mat).tocsr()


@dataclass
class SyntheticData:
    name: str
    n_users: int
    n_items: int
    embed_size: int
    n_layers: int
    dv: int
    dt: int
    train: sp.csr_matrix          # raw interactions (U x I, ones)
    ui_norm: sp.csr_matrix        # D^-1/2 R
    iu_norm: sp.csr_matrix        # D^-1/2 R^T

    @property
    def nnz(self) -> int:
        return int(self.train.nnz)


def make_dataset(name: str, seed: int = 2022, uniform: bool = False) -> SyntheticData:
    U, I, nnz, d, k, dv, dt = CONFIGS[name]
    r = make_bipartite(U, I, nnz, seed=seed, uniform=uniform)
    return SyntheticData(name, U, I, d, k, dv, dt, r, csr_norm(r), csr_norm(r.T.tocsr()))


def make_features(n_items: int, dim: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal((n_items, dim), dtype=np.float32)


class TripleSampler:
    """B distinct users with >= 1 interaction (when B <= #such users), one uniform positive from the
    user's row, one uniform negative by rejection against the row (load_data.py:153-180)."""

    def __init__(self, train: sp.csr_matrix, seed: int = 2022):
        self.indptr, self.indices = train.indptr.astype(np.int64), train.indices.astype(np.int64)
        self.n_users, self.n_items = train.shape
        self.deg = np.diff(self.indptr)
        self.exist = np.nonzero(self.deg > 0)[0]
        self.rng = np.random.default_rng(seed)
        # sorted (user, item) keys for vectorised membership tests
        users = np.repeat(np.arange(self.n_users, dtype=np.int64), self.deg)
        self.keys = np.sort(users * self.n_items + self.indices)

    def sample(self, batch: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        rng = self.rng
        users = rng.choice(self.exist, size=batch, replace=batch > len(self.exist))
        off = (rng.random(batch) * self.deg[users]).astype(np.int64)
        pos = self.indices[self.indptr[users] + off]
        neg = rng.integers(0, self.n_items, batch)
        for _ in range(1000):
            k = users * self.n_items + neg
            j = np.searchsorted(self.keys, k)
            bad = (j < len(self.keys)) & (self.keys[np.minimum(j, len(self.keys) - 1)] == k)
            if not bad.any():
                break
            neg[bad] = rng.integers(0, self.n_items, int(bad.sum()))
        return users.astype(np.int64), pos.astype(np.int64), neg.astype(np.int64)
