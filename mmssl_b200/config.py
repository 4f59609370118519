"""Hot-path hyper-parameters with the reference's defaults (utility/parser.py of the reference:
:54-55 batch/embed, :60 cl_rate, :64 regs, :65 lr, :72-77 GNN rates, :82 weight_size, :94 tau,
:29 feat_reg_decay, :12 layers, :14 sparse).  ``default_args()`` returns an argparse-like Namespace
so that ``Models.py`` can keep the reference's module-global ``args`` convention when it is used
outside the reference tree."""
from __future__ import annotations

import argparse


def default_args(**overrides) -> argparse.Namespace:
    ns = argparse.Namespace(
        embed_size=64, batch_size=1024, weight_size="[64, 64]", head_num=4, layers=1, sparse=1,
        drop_rate=0.2, model_cat_rate=0.55, id_cat_rate=0.36, tau=0.5, cl_rate=0.03, regs="[1e-5,1e-5,1e-2]",
        feat_reg_decay=1e-5, lr=0.00055, G_drop1=0.31, G_drop2=0.5, seed=2022,
    )
    for k, v in overrides.items():
        setattr(ns, k, v)
    return ns
