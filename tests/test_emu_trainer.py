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
