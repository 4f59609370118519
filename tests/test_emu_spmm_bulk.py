"""The bulk-copy SpMM (csrc/spmm_bulk.cu) and its plan executed on the CPU under the cuemu emulator (fibers + the functional
model of mbarrier / cp.async.bulk): the bodies of tests/test_gpu_spmm_bulk.py, unchanged."""
import pytest

from tests.cuemu import harness


@pytest.fixture
def emu(monkeypatch):
    harness.set_order("fwd")
    return harness.emulated_device(monkeypatch)


def test_plan(emu):
    from tests import test_gpu_spmm_bulk as T
    T.test_bulk_plan_covers_every_nonzero_once()


@pytest.mark.parametrize("d,nrhs,variant", [(64, 1, (0, 0, 0, 0)), (64, 2, (0, 2, 3, 0)), (128, 1, (0, 4, 1, 1)), (256, 2, (0, 1, 2, 0)), (128, 2, (0, 0, 0, 1)), (64, 1, (0, 8, 1, 1)), (256, 1, (0, 0, 0, 0))])
def test_plain(emu, d, nrhs, variant):
    from tests import test_gpu_spmm_bulk as T
    T.test_spmm_bulk_plain(d, nrhs, variant)


@pytest.mark.parametrize("d,tma", [(64, 0), (64, 1), (128, 0), (256, 0), (128, 1)])
def test_epilogues(emu, d, tma):
    from tests import test_gpu_spmm_bulk as T
    T.test_spmm_bulk_epilogues(d, tma)


def test_short_empty_rows_empty_graph_heavy_rows(emu):
    from tests import test_gpu_spmm_bulk as T
    T.test_spmm_bulk_many_short_and_empty_rows()
    T.test_spmm_bulk_empty_graph()
    T.test_spmm_bulk_heavy_rows_and_zipf_columns()
