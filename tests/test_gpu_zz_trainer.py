"""mmssl_b200.trainer on the GPU: the reference's Trainer life cycle (model construction, full steps, evaluation, best-recall /
early stopping) on the small reference-format dataset.  Same body as tests/test_emu_trainer.py."""
import pytest

from tests import trainer_check

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sampler", ["reference", "device"])
def test_trainer_runs_the_reference_life_cycle(sampler):
    trainer_check.run_life_cycle("cuda", sampler)
