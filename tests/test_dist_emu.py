"""The row-sharded whole hot step (mmssl_b200/rowshard_step.py) with world_size 2 over gloo, every rank running
the REAL kernel sources under the cuemu emulator: losses, the rank's rows of both table gradients, the all-reduced small
gradients and the parameters after two AdamW steps against the single-process fused HotStep on the same problem."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

WORLD = 2


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class _MP:
    def setattr(self, o, n, v):
        setattr(o, n, v)


def _problem(modal):
    import scipy.sparse as sp
    from mmssl_b200.synthetic import csr_norm, make_bipartite
    U, I, d, B = 203, 131, 64, 48                         # odd sizes: the last rank's blocks are padded
    r = make_bipartite(U, I, 1500, seed=5)
    g = torch.Generator().manual_seed(2)
    xav = lambda a, b: (torch.rand(a, b, generator=g) * 2 - 1) * (6.0 / (a + b)) ** 0.5
    P = {"image_trans.weight": xav(d, 40), "image_trans.bias": torch.randn(d, generator=g) * 0.1, "text_trans.weight": xav(d, 24),
         "text_trans.bias": torch.randn(d, generator=g) * 0.1, "user_id_embedding.weight": xav(U, d), "item_id_embedding.weight": xav(I, d),
         "weight_dict.w_self_attention_cat": xav(4 * d, d)}
    feats = (torch.randn(I, 40, generator=g), torch.randn(I, 24, generator=g))
    masks = tuple(((torch.rand(I, d, generator=g) >= 0.2) / 0.8).float() for _ in range(2))
    users = torch.randperm(U, generator=g)[:B]
    pos, neg = torch.randint(0, I, (B,), generator=g), torch.randint(0, I, (B,), generator=g)
    mods = None
    if modal == "random":                                 # distinct image / text graphs with duplicate entries
        rng = np.random.default_rng(3)
        mk = lambda nnz: sp.csr_matrix((np.ones(nnz, np.float32), (rng.integers(0, U, nnz), rng.integers(0, I, nnz))), shape=(U, I))
        mods = [mk(700), mk(400)]
        mods = [(csr_norm(m), csr_norm(m.T.tocsr())) for m in mods]
    return U, I, d, B, csr_norm(r), csr_norm(r.T.tocsr()), P, feats, masks, (users, pos, neg), mods


def _worker(rank, port, modal, schedule, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from tests.cuemu import harness
        harness.set_order("fwd")
        harness.emulated_device(_MP())
        from mmssl_b200.engine import LIVE, P_EI, P_EU, FeatureStore
        from mmssl_b200.graph import BipartiteGraph
        from mmssl_b200.hotstep import HotStep, HotStepConfig
        from mmssl_b200.rowshard_step import RowBlockGraph, RowShardedHotStep, shard_problem
        from tests.golden_util import rel_err
        U, I, d, B, a_ui, a_iu, P, feats, masks, (users, pos, neg), mods = _problem(modal)
        cfg = HotStepConfig(embed_size=d, n_layers=2, batch_size=B, proj_impl="simt")
        # ---- single-process reference: the fused HotStep on the full problem
        g_ui, g_iu = BipartiteGraph.from_scipy(a_ui, device="cpu"), BipartiteGraph.from_scipy(a_iu, device="cpu")
        graphs = [g_ui, g_iu, g_ui, g_iu, g_ui, g_iu]
        if mods is not None:
            for j, (m_ui, m_iu) in zip((2, 4), mods):
                graphs[j], graphs[j + 1] = BipartiteGraph.from_scipy(m_ui, device="cpu"), BipartiteGraph.from_scipy(m_iu, device="cpu")
        Pf = {k: v.clone() for k, v in P.items()}
        hs = HotStep(Pf, tuple(FeatureStore(f.clone()) for f in feats), graphs, cfg, batch=B)
        hs.engine.two_streams = False
        hs.masks = masks
        hs.set_indices(users, pos, neg)
        # ---- this rank of the sharded step
        Pl, fl, gl, pu, pi = shard_problem(P, feats, a_ui, a_iu, rank, WORLD, "cpu")
        gl = list(gl)
        if mods is not None:
            for j, (m_ui, m_iu) in zip((2, 4), mods):
                gl[j], gl[j + 1] = RowBlockGraph.from_scipy(m_ui, pu, pi, rank, "cpu"), RowBlockGraph.from_scipy(m_iu, pi, pu, rank, "cpu")
        sh = RowShardedHotStep(Pl, fl, gl, cfg, B, pu, pi, rank, schedule=schedule)
        sh.masks = tuple(pi.local(m, rank) for m in masks)
        sh.set_indices(users, pos, neg)
        errs = {}
        for step in range(2):
            want = hs.run().clone()
            got = sh.run().clone()
            errs[f"loss{step}"] = float(((got - want).abs() / want.abs().clamp_min(1e-12)).max())
            ulo, uhi = pu.bounds(rank)
            ilo, ihi = pi.bounds(rank)
            for k in LIVE:
                w = hs.grads[k]
                if k == P_EU:
                    errs[f"g{step}/{k}"] = rel_err(sh.grads[k][:uhi - ulo], w[ulo:uhi])
                elif k == P_EI:
                    errs[f"g{step}/{k}"] = rel_err(sh.grads[k][:ihi - ilo], w[ilo:ihi])
                else:
                    errs[f"g{step}/{k}"] = rel_err(sh.grads[k], w)
        for k in LIVE:
            if k == P_EU:
                errs["p/" + k] = rel_err(sh.P[k][:uhi - ulo], hs.P[k][ulo:uhi])
            elif k == P_EI:
                errs["p/" + k] = rel_err(sh.P[k][:ihi - ilo], hs.P[k][ilo:ihi])
            else:
                errs["p/" + k] = rel_err(sh.P[k], hs.P[k])
        errs["gathers_per_step"] = sh.n_gathers / 2
        errs["reduce_scatters_per_step"] = sh.n_reduce_scatters / 2
        ret[rank] = errs
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("modal,schedule", [("alias", "reduce_scatter"), ("random", "reduce_scatter"), ("alias", "allgather"), ("random", "allgather")])
def test_row_sharded_hot_step_matches_single_process(modal, schedule):
    """schedule 'allgather': every product all-gathers its dense operand; 'reduce_scatter': products whose operand is user-sized
    multiply the rank's column block and reduce-scatter the item-sized result (mmssl_reduce_rows_epilogue applies the epilogue)."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(port, modal, schedule, ret), nprocs=WORLD, join=True)
    assert len(ret) == WORLD
    for rank in range(WORLD):
        e = dict(ret[rank])
        gathers, rs = e.pop("gathers_per_step"), e.pop("reduce_scatters_per_step")
        bad = {k: v for k, v in e.items() if not v < 2e-5}
        assert not bad, (rank, bad)
        assert gathers > 0 and (rs > 0) == (schedule == "reduce_scatter")


def _disk_worker(rank, port, shard_dir, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from tests.cuemu import harness
        harness.set_order("fwd")
        harness.emulated_device(_MP())
        from mmssl_b200.dataset import ReferenceDataset
        from mmssl_b200.engine import LIVE
        from mmssl_b200.hotstep import HotStepConfig
        from mmssl_b200.rowshard_step import RowShardedHotStep, shard_problem, shard_problem_from_disk
        from mmssl_b200.synthetic import csr_norm
        from tests.golden_util import rel_err
        ds = ReferenceDataset.load(os.path.join(os.path.dirname(__file__), "golden", "dataset_small"))
        R = ds.train_mat.astype(np.float32).tocsr()
        U, I = R.shape
        d, B = 64, 16
        g = torch.Generator().manual_seed(4)
        xav = lambda a, b: (torch.rand(a, b, generator=g) * 2 - 1) * (6.0 / (a + b)) ** 0.5
        P = {"image_trans.weight": xav(d, ds.image_feats.shape[1]), "image_trans.bias": torch.zeros(d),
             "text_trans.weight": xav(d, ds.text_feats.shape[1]), "text_trans.bias": torch.zeros(d),
             "user_id_embedding.weight": xav(U, d), "item_id_embedding.weight": xav(I, d), "weight_dict.w_self_attention_cat": xav(4 * d, d)}
        users = torch.randperm(U, generator=g)[:B]
        pos, neg = torch.randint(0, I, (B,), generator=g), torch.randint(0, I, (B,), generator=g)
        cfg = HotStepConfig(embed_size=d, n_layers=2, batch_size=B, drop_rate=0.0, proj_impl="simt")
        feats = (torch.from_numpy(np.asarray(ds.image_feats, np.float32)), torch.from_numpy(np.asarray(ds.text_feats, np.float32)))
        steps = []
        for build in (lambda: shard_problem(P, feats, csr_norm(R), csr_norm(R.T.tocsr()), rank, WORLD, "cpu"),
                      lambda: shard_problem_from_disk(shard_dir, P, rank, WORLD, "cpu")):
            Pl, fl, gl, pu, pi = build()
            sh = RowShardedHotStep(Pl, fl, gl, cfg, B, pu, pi, rank)
            sh.set_indices(users, pos, neg)
            out = sh.run().clone()
            steps.append((out, {k: sh.P[k].clone() for k in LIVE}))
        (o_mem, p_mem), (o_disk, p_disk) = steps
        ret[rank] = dict(loss=float((o_mem - o_disk).abs().max()), params=max(rel_err(p_disk[k], p_mem[k]) for k in LIVE))
    finally:
        dist.destroy_process_group()


def test_row_sharded_step_from_memory_mapped_shards(tmp_path):
    """dataset.write_shards -> ShardedDataset (every rank maps only its row blocks) -> RowShardedHotStep: identical to the step
    built from the full in-memory matrices."""
    from mmssl_b200.dataset import ReferenceDataset, write_shards
    ds = ReferenceDataset.load(os.path.join(os.path.dirname(__file__), "golden", "dataset_small"))
    write_shards(ds, str(tmp_path))
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_disk_worker, args=(port, str(tmp_path), ret), nprocs=WORLD, join=True)
    for rank in range(WORLD):
        assert ret[rank]["loss"] == 0.0 and ret[rank]["params"] == 0.0, dict(ret[rank])
