"""On-disk formats: the loader against what the unmodified reference loader produced on the same
directory (tests/golden/dataset_small/expected.json, minted by make_golden_dataset.py), and the memory-mapped per-rank
shards against scipy row slicing."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from mmssl_b200.dataset import OPERANDS, ROW_SPACE, ReferenceDataset, ShardedDataset, write_shards
from mmssl_b200.parallel import RowPartition, shard_rows_scipy
from mmssl_b200.synthetic import csr_norm

ROOT = os.path.join(os.path.dirname(__file__), "golden", "dataset_small")


def test_loader_matches_reference_data_class():
    exp = json.load(open(os.path.join(ROOT, "expected.json")))
    ds = ReferenceDataset.load(ROOT)
    for k in ("n_users", "n_items", "n_train", "n_test", "exist_users"):
        assert getattr(ds, k) == exp[k], k
    for k in ("train_items", "test_set", "val_set"):
        assert {str(u): v for u, v in getattr(ds, k).items()} == exp[k], k
    assert 5 not in ds.train_items and 6 not in ds.test_set          # empty lists are dropped (load_data.py:64-66, :74-76)
    assert ds.train_mat.shape == (61, 43) and ds.image_feats.shape == (43, 12) and ds.text_feats.shape == (43, 8)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shards_roundtrip_per_rank(tmp_path, world):
    ds = ReferenceDataset.load(ROOT)
    meta = write_shards(ds, str(tmp_path))
    R = ds.train_mat.astype(np.float32).tocsr()
    a_ui, a_iu = csr_norm(R), csr_norm(R.T.tocsr())
    full = {"ui": a_ui, "iu": a_iu, "iuT": a_iu.T.tocsr(), "uiT": a_ui.T.tocsr()}
    assert meta["n_users"] == 61 and meta["n_items"] == 43
    feats = {"image": [], "text": []}
    touched = 0
    for rank in range(world):
        sh = ShardedDataset.open(str(tmp_path), rank, world)
        for name in OPERANDS:
            part = RowPartition(full[name].shape[0], world)
            want = shard_rows_scipy(full[name], part, rank)
            blk = sh.operand(name)
            got = blk.to_scipy()
            assert got.shape == want.shape == (part.block, full[name].shape[1])
            assert (got != want).nnz == 0
            assert blk.indices.dtype == np.int32 and blk.values.dtype == np.float32 and blk.indptr.dtype == np.int64
            assert ROW_SPACE[name] in ("user", "item") and (blk.lo, blk.hi) == part.bounds(rank)
        for which in feats:
            feats[which].append(np.asarray(sh.features(which)))
        touched += sh.bytes_touched()
    assert np.array_equal(np.concatenate(feats["image"]), ds.image_feats)
    assert np.array_equal(np.concatenate(feats["text"]), ds.text_feats)
    one = ShardedDataset.open(str(tmp_path), 0, 1).bytes_touched()
    assert touched <= one + 8 * 4 * world * (RowPartition(61, world).block + 1)      # ranks share nothing but padded indptr


def test_shard_directory_is_validated(tmp_path):
    (tmp_path / "meta.json").write_text(json.dumps({"format": "something else"}))
    with pytest.raises(ValueError):
        ShardedDataset.open(str(tmp_path))
