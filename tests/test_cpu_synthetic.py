import numpy as np

from mmssl_b200.synthetic import TripleSampler, csr_norm, make_bipartite, make_dataset
from oracle import mmssl_oracle as O


def test_bipartite_exact_nnz_and_norm():
    r = make_bipartite(500, 300, 4000, seed=1)
    assert r.nnz == 4000 and r.data.max() == 1.0
    assert (np.diff(r.indptr) >= 1).all()
    a = csr_norm(r)
    b = O.csr_norm(r, mean_flag=True)
    assert abs(a - b).max() < 1e-12


def test_sampler_semantics():
    ds = make_dataset("tiny")
    s = TripleSampler(ds.train, seed=0)
    u, p, n = s.sample(128)
    assert len(set(u.tolist())) == 128                      # distinct users (load_data.py:154-155)
    dense = ds.train.toarray()
    assert (dense[u, p] == 1).all() and (dense[u, n] == 0).all()
