"""mmssl_b200.trainer: the reference's Trainer life cycle (main.py:31-498) over the device-resident full step.  The host
sampler is pinned to batches recorded from the unmodified reference's ``Data.sample``; the epoch loop runs end to end on
the small reference-format dataset under the cuemu emulator (model construction, full steps, evaluation, early stopping)."""
import json
import os
import random

import numpy as np
import pytest

from mmssl_b200.dataset import ReferenceDataset
from tests import trainer_check
from tests.cuemu import harness

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("batch", [16, 80])
def test_reference_sampler_draws_the_reference_batches(batch):
    from mmssl_b200.trainer import reference_sample
    want = json.load(open(os.path.join(GOLD, "sampler_small.json")))[str(batch)]
    ds = ReferenceDataset.load(os.path.join(GOLD, "dataset_small"))
    random.seed(11)
    np.random.seed(11)
    for w in want:
        users, pos, neg = reference_sample(ds, batch)
        assert [list(users), list(pos), list(neg)] == w


@pytest.mark.parametrize("sampler", ["reference", "device"])
def test_trainer_runs_the_reference_life_cycle(monkeypatch, sampler):
    harness.set_order("fwd")
    # red zones around every tensor: the fixture has an item id beyond train_mat's width, which must never reach a kernel
    harness.emulated_device(monkeypatch, guard=(sampler == "reference"))
    trainer_check.run_life_cycle("cpu", sampler)


def test_epoch_loop_bookkeeping_matches_main_py():
    """Best-recall / early-stopping rule and evaluation cadence of main.py:436-498 on a scripted sequence of validation recalls
    (no kernels involved): test() on the test split only when the validation recall improves; `patience` non-improving epochs are
    tolerated, the next one stops the run."""
    import torch
    from mmssl_b200.trainer import Trainer, TrainerArgs

    class Data:
        n_train, val_set, test_set = 40, {0: [1], 1: [2]}, {0: [3], 5: [4]}

    class Step:
        idx, epochs = 0, 0

        def start_epoch(self):
            self.idx, self.epochs = 0, self.epochs + 1

        def step(self, u, p, n):
            self.idx += 1
            z = torch.zeros(())
            return dict(batch_loss=z + 1.0, loss5=torch.tensor([1.0, 0.5, 0.25, 0.0, 0.125]))

    recalls = [0.10, 0.20, 0.15, 0.20, 0.30, 0.10, 0.10, 0.10, 0.50]
    calls = []
    tr = object.__new__(Trainer)
    tr.args = TrainerArgs(epoch=20, batch_size=16, verbose=1, early_stopping_patience=2)
    tr.data, tr.device, tr.step, tr.Ks, tr.cuda_graph = Data(), torch.device("cpu"), Step(), [10, 20, 50], False
    lines = []
    tr.log = lines.append
    tr.sample = lambda: ([0], [0], [0])

    def fake_test(users, is_val):
        calls.append(("val" if is_val else "test", sorted(users)))
        r = recalls[tr.step.epochs - 1] if is_val else 0.9
        a = np.array([r / 2, r, r * 2])
        return {"recall": a, "precision": a, "ndcg": a, "hit_ratio": a, "auc": 0.}
    tr.test = fake_test
    best, test_ret = tr.train()
    # epochs 1..5 improve at 1, 2, 5; epochs 6, 7 are tolerated (patience 2), epoch 8 stops: the 0.50 of epoch 9 is never seen
    assert tr.step.epochs == 8 and best == pytest.approx(0.30)
    assert [c[0] for c in calls] == ["val", "test", "val", "test", "val", "val", "val", "test", "val", "val", "val"]
    assert calls[0][1] == [0, 1] and calls[1][1] == [0, 5]                       # validation users / test users (main.py:449-451)
    assert tr.step.idx == Data.n_train // 16 + 1                                   # batches per epoch (main.py:328)
    assert sum(l.startswith("#####Early stopping steps") for l in lines) == 4 and lines[-2] == "#####Early stop! #####"
    assert test_ret["recall"][1] == pytest.approx(0.9)
    assert lines[0].startswith("Epoch 0 [") and "train==[3.00000=1.50000 + 0.75000 + 0.00000]" in lines[0]
