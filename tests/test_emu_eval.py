"""The evaluation kernel (csrc/eval.cu: fused score -> mask -> top-K -> metrics) executed on the CPU by the cuemu
fiber emulator through mmssl_b200/evaluate.py -- the bodies of tests/test_gpu_zz_eval.py, unchanged: exact ranking
(incl. tie order), hit lists and fp64 metrics against the oracle and the golden vectors minted from the reference."""
import functools

import numpy as np
import pytest

from tests import test_gpu_zz_eval as E
from tests.cuemu import harness


@pytest.fixture(params=["fwd", "rev"])
def emu(request, monkeypatch):
    harness.set_order(request.param)
    lib = harness.emulated_device(monkeypatch)
    from mmssl_b200 import evaluate
    monkeypatch.setattr(evaluate, "Evaluator", functools.partial(evaluate.Evaluator, device="cpu"))
    return lib


@pytest.mark.parametrize("case", ["eval_random", "eval_ties", "eval_short"])
@pytest.mark.parametrize("split", ["test", "val"])
def test_eval_matches_reference_golden(emu, case, split):
    E.test_eval_matches_reference_golden(case, split)


def test_eval_many_compactions(emu):
    """More items than the candidate buffer holds, increasing scores along the item axis (worst case for the
    threshold filter: every sweep appends), users not a multiple of the 8-user tile, an empty training row."""
    from mmssl_b200.evaluate import Evaluator
    rng = np.random.default_rng(5)
    U, I, d, Ks = 21, 3000, 8, [5, 20, 64]
    ua = np.abs(rng.standard_normal((U, d))).astype(np.float32)
    ia = (np.abs(rng.standard_normal((I, d))) * np.linspace(0.1, 3.0, I)[:, None]).astype(np.float32)
    train = {u: sorted(rng.choice(I, size=int(rng.integers(0, 400)), replace=False).tolist()) for u in range(U)}
    train[3] = []
    held = {u: rng.choice(I, size=int(rng.integers(1, 30)), replace=False).tolist() for u in range(U)}
    ev = Evaluator({u: v for u, v in train.items() if v}, held, {}, U, I, Ks)

    def csr(rows):
        ptr = np.zeros(U + 1, np.int64)
        for u in range(U):
            ptr[u + 1] = ptr[u] + len(rows.get(u, []))
        idx = np.concatenate([np.sort(np.asarray(rows.get(u, []), np.int64)) for u in range(U)]) if ptr[-1] else np.zeros(0, np.int64)
        return ptr, idx

    users = np.arange(U, dtype=np.int64)[::-1].copy()
    E._check_against_oracle(ev, ua, ia, users, csr(train), csr(held), Ks, False)


def test_eval_random_tie_heavy_cases(emu):
    """Small-integer embeddings (many exactly equal scores), random sizes, widths, cut-offs, training rows from empty to almost
    everything: ranking incl. tie order, hit lists and fp64 metrics exactly equal to the oracle."""
    from mmssl_b200.evaluate import Evaluator
    rng = np.random.default_rng(123)
    for case in range(12):
        U, I = int(rng.integers(1, 40)), int(rng.integers(1, 700))
        d = int(rng.choice([4, 8, 64, 128]))
        Ks = sorted(set(int(k) for k in rng.integers(1, 65, int(rng.integers(1, 4)))))
        ua = rng.integers(-2, 3, (U, d)).astype(np.float32)
        ia = rng.integers(-2, 3, (I, d)).astype(np.float32)
        train = {u: sorted(rng.choice(I, size=int(rng.integers(0, I)), replace=False).tolist()) for u in range(U)}
        held = {u: rng.choice(I, size=int(rng.integers(1, min(6, I) + 1)), replace=False).tolist() for u in range(U)}
        ev = Evaluator({u: v for u, v in train.items() if v}, held, {}, U, I, Ks)

        def csr(rows):
            ptr = np.zeros(U + 1, np.int64)
            for u in range(U):
                ptr[u + 1] = ptr[u] + len(rows.get(u, []))
            idx = np.concatenate([np.sort(np.asarray(rows.get(u, []), np.int64)) for u in range(U)]) if ptr[-1] else np.zeros(0, np.int64)
            return ptr, idx
        E._check_against_oracle(ev, ua, ia, rng.permutation(U).astype(np.int64), csr(train), csr(held), Ks, False)
