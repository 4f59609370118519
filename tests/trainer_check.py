"""Shared body of the Trainer life-cycle tests (CPU emulation: tests/test_emu_trainer.py, GPU: tests/test_gpu_zz_trainer.py)."""
import os

import numpy as np

from mmssl_b200.dataset import ReferenceDataset

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def run_life_cycle(device: str, sampler: str, epochs: int = 3):
    from mmssl_b200.trainer import Trainer, TrainerArgs, set_seed
    ds = ReferenceDataset.load(os.path.join(GOLD, "dataset_small"))
    args = TrainerArgs(dataset="dataset_small", epoch=epochs, batch_size=16, verbose=2, early_stopping_patience=1, m_topk_rate=0.05,
                       Ks="[2, 5, 10]", seed=5)
    set_seed(args.seed)
    lines = []
    tr = Trainer(ds, args, device=device, sampler=sampler, log=lines.append)
    assert tr.n_users == 61 and tr.n_items == 43 and tr.step.k == 2
    keys0 = set(tr.model.state_dict().keys())
    p0 = tr.model.user_id_embedding.weight.detach().clone()
    d0 = tr.D.net[0].weight.detach().clone()
    best, test_ret = tr.train()
    n_batch = ds.n_train // 16 + 1
    assert tr.step.idx == n_batch                                   # the iteration counter restarts every epoch
    assert len(tr.history) >= 2 and all(np.isfinite(h["loss"]) for h in tr.history)
    # the nn.Modules are the checkpoint: FullStep updated their parameters in place, names unchanged
    assert set(tr.model.state_dict().keys()) == keys0
    assert float((tr.model.user_id_embedding.weight.detach() - p0).abs().max()) > 0
    assert float((tr.D.net[0].weight.detach() - d0).abs().max()) > 0
    assert int(tr.D.net[2].num_batches_tracked) == 4 * n_batch * len(tr.history)      # 4 D calls per iteration
    assert test_ret is not None and set(test_ret) == {"precision", "recall", "ndcg", "hit_ratio", "auc"}
    assert len(test_ret["recall"]) == 3 and 0.0 <= best <= 1.0
    assert any(l.startswith("Epoch 0 [") for l in lines) and any(l.startswith("Test_Recall@5") for l in lines)
    ret = tr.test(list(ds.val_set.keys()), is_val=True)
    assert abs(float(ret["recall"][1]) - tr.history[-1]["recall"]) < 1e-12   # test() is deterministic (eval mode)
    return tr
