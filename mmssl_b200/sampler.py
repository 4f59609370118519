"""Device-side triple sampler (SURVEY 8f "next" #1): the reference's ``Data.sample``
(utility/load_data.py:153-191) as one CUDA kernel, so a training step needs no host input at all."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import ptr, stream


class DeviceTripleSampler:
    def __init__(self, train_csr, device="cuda", seed: int = 2022):
        lib = _lib.load(require_device=True)
        csr = train_csr.tocsr()
        csr.sort_indices()
        self.n_users, self.n_items = csr.shape
        self.indptr = torch.from_numpy(csr.indptr.astype(np.int64)).to(device)
        self.indices = torch.from_numpy(csr.indices.astype(np.int64)).to(device)
        exist = np.nonzero(np.diff(csr.indptr) > 0)[0].astype(np.int64)
        self.exist = torch.from_numpy(exist).to(device)
        self.claim = torch.empty(max(len(exist), 1), dtype=torch.int32, device=device)
        self.seed = seed
        _lib.check(lib.mmssl_sampler_init(ptr(self.claim), len(exist), stream()))

    def sample_into(self, out: torch.Tensor, step_dev: torch.Tensor = None, step: int = 0) -> torch.Tensor:
        """Fills out[3, B] (int64: users, pos, neg).  `step_dev` (int32 device scalar) makes the launch
        replayable inside a CUDA graph with a fresh batch per replay."""
        lib = _lib.load(require_device=True)
        assert out.dtype == torch.int64 and out.dim() == 2 and out.shape[0] == 3 and out.is_contiguous()
        b = out.shape[1]
        _lib.check(lib.mmssl_sample_triples(ptr(self.indptr), ptr(self.indices), ptr(self.exist), self.exist.numel(),
                                            self.n_items, b, self.seed, ptr(step_dev), int(step), ptr(self.claim),
                                            ptr(out[0]), ptr(out[1]), ptr(out[2]), stream()))
        return out
