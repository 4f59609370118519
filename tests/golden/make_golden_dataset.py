#!/usr/bin/env python
"""Writes a small dataset directory in the reference's on-disk format (tests/golden/dataset_small/) and records what the
UNMODIFIED reference loader (utility/load_data.py:Data) makes of it -> expected.json.  Build-container only."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, make_dataset  # noqa: E402


def main():
    name = "dataset_small"
    make_dataset(HERE, name, 61, 43, 12, 8, seed=21)
    d = os.path.join(HERE, name)
    # irregularities the loader has to survive: an empty training list, an empty held-out list, a val-only item id
    tr = json.load(open(os.path.join(d, "train.json")))
    te = json.load(open(os.path.join(d, "test.json")))
    va = json.load(open(os.path.join(d, "val.json")))
    tr["5"] = []
    te["6"] = []
    va["9"] = [44]                                    # larger than any train / test id -> n_items = 45
    for fn, obj in (("train.json", tr), ("test.json", te), ("val.json", va)):
        json.dump(obj, open(os.path.join(d, fn), "w"))
    sys.path.insert(0, REF)
    os.chdir(REF)
    sys.argv = ["x", "--debug"]
    from utility.load_data import Data
    dg = Data(path=d, batch_size=16)
    exp = dict(n_users=dg.n_users, n_items=dg.n_items, n_train=dg.n_train, n_test=dg.n_test, exist_users=dg.exist_users,
               train_items={str(k): v for k, v in dg.train_items.items()}, test_set={str(k): v for k, v in dg.test_set.items()},
               val_set={str(k): v for k, v in dg.val_set.items()})
    json.dump(exp, open(os.path.join(d, "expected.json"), "w"))
    print("wrote", d, {k: exp[k] for k in ("n_users", "n_items", "n_train", "n_test")})


if __name__ == "__main__":
    main()
