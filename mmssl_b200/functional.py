"""torch.autograd.Function wrappers: the boundary between the reference's PyTorch training loop and
the CUDA library.  Tensors in, tensors out, gradients by hand-written backward kernels.

  SpMMFunction            <- MMSSL.mm / torch.sparse.mm            (Models.py:69-73, :203-208)
  MMSSLForwardFn          <- MMSSL.forward as one node             (Models.py:171-220)
  bpr_loss                <- Trainer.bpr_loss                      (main.py:499-511)
  batched_contrastive_loss<- Trainer.batched_contrastive_loss      (main.py:218-249)
  feat_reg_loss           <- Trainer.feat_reg_loss_calculation     (main.py:252-257)
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops
from .engine import LIVE, Engine
from .graph import BipartiteGraph, prepare


class SpMMFunction(torch.autograd.Function):
    """Y = A @ X with dX = A^T @ dY; the graph never requires grad (Models.py:69-73)."""

    @staticmethod
    def forward(ctx, graph: BipartiteGraph, x: torch.Tensor):
        ctx.graph = graph
        xc = x if (x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0) else x.contiguous()
        if graph.nnz == 0:
            return torch.zeros(graph.shape[0], x.shape[1], dtype=torch.float32, device=x.device)
        return ops.spmm(graph.fwd, [xc.detach()])[0]

    @staticmethod
    def backward(ctx, gy):
        g = ctx.graph
        if g.nnz == 0:
            return None, torch.zeros(g.shape[1], gy.shape[1], dtype=torch.float32, device=gy.device)
        return None, ops.spmm(g.bwd, [gy.contiguous()])[0]


def spmm(graph, x: torch.Tensor) -> torch.Tensor:
    """Drop-in for ``torch.sparse.mm(graph, x)``: `graph` is a torch sparse COO tensor or a prepared graph."""
    return SpMMFunction.apply(prepare(graph), x)


class MMSSLForwardFn(torch.autograd.Function):
    """The whole ``MMSSL.forward`` as a single autograd node over the live parameters."""

    @staticmethod
    def forward(ctx, engine: Engine, feats, graphs, masks, w_v, b_v, w_t, b_t, e_u, e_i, w_cat):
        P = dict(zip(LIVE, (w_v.detach(), b_v.detach(), w_t.detach(), b_t.detach(), e_u.detach(), e_i.detach(), w_cat.detach())))
        outs, st = engine.forward(P, feats, graphs, masks, want_sumsq=False)
        ctx.engine, ctx.st, ctx.P, ctx.feats = engine, st, P, feats
        outs = list(outs)
        # autograd wants distinct tensor objects per output
        if outs[7] is outs[6]:
            outs[7] = outs[6].view_as(outs[6])
        if outs[9] is outs[8]:
            outs[9] = outs[8].view_as(outs[8])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        res = ctx.engine.backward(ctx.st, ctx.P, ctx.feats, grads)
        return (None, None, None, None) + tuple(res[k] for k in LIVE)


# --------------------------------------------------------------------------------------------------
class _BprFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, p, n, reg_coef: float):
        u, p, n = (t.contiguous() for t in (u, p, n))
        part, nb = ops.bpr(u, p, n, None, None, None, mode=1, reg_coef=reg_coef)
        out5 = torch.empty(5, dtype=torch.float32, device=u.device)
        ops.loss_assemble(part, nb, u.shape[0], reg_coef, None, None, 0.0, None, None, 0, 0.0, out5)
        ctx.save_for_backward(u, p, n)
        ctx.reg_coef = reg_coef
        return out5[1], out5[2]

    @staticmethod
    def backward(ctx, g_mf, g_emb):
        u, p, n = ctx.saved_tensors
        gu, gp, gn = torch.zeros_like(u), torch.zeros_like(p), torch.zeros_like(n)
        g_mf = g_mf.contiguous().float()
        g_emb = g_emb.contiguous().float()
        ops.bpr(u, p, n, None, None, None, mode=2, reg_coef=ctx.reg_coef, g_mf=g_mf, g_emb=g_emb, g_u=gu, g_p=gp, g_n=gn)
        return gu, gp, gn, None


def bpr_loss(users_emb, pos_emb, neg_emb, decay: float = 1e-5, batch_size: int = 1024):
    """(mf_loss, emb_loss, reg_loss) exactly as Trainer.bpr_loss (main.py:499-511); the regulariser
    divides by the configured batch size (main.py:504)."""
    mf, emb = _BprFn.apply(users_emb, pos_emb, neg_emb, float(decay) / float(batch_size))
    return mf, emb, 0.0


class _InfoNCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, tau: float):
        z1, z2 = z1.contiguous(), z2.contiguous()
        n, d = z1.shape
        work = ops.InfoNCEWork(n, d, z1.device)
        part = ops.infonce_forward(z1, z2, None, 1.0 / tau, work)
        out5 = torch.empty(5, dtype=torch.float32, device=z1.device)
        ops.loss_assemble(None, 0, 0, 0.0, None, None, 0.0, part, None, n, 1.0, out5)
        ctx.work, ctx.tau = work, tau
        ctx.save_for_backward(z1, z2)
        return out5[4]

    @staticmethod
    def backward(ctx, g):
        z1, z2 = ctx.saved_tensors
        w = ctx.work
        # coefficients with the actual upstream gradient as seed (device scalar, no host sync)
        ops.infonce_forward(z1, z2, None, 1.0 / ctx.tau, w, g_loss=g.contiguous().float())
        g1, g2 = torch.zeros_like(z1), torch.zeros_like(z2)
        ops.infonce_backward(None, 1.0 / ctx.tau, w, g1, g2)
        return g1, g2, None


def batched_contrastive_loss(z1, z2, tau: float = 0.5, batch_size: int = 1024):
    """Trainer.batched_contrastive_loss (main.py:218-249).  The reference's block loops concatenate
    over all column blocks, so the value does not depend on `batch_size`; one fused pass here."""
    return _InfoNCEFn.apply(z1, z2, float(tau))


class _FeatRegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coef: float, *xs):
        xs = [x if (x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0) else x.contiguous() for x in xs]
        parts = torch.cat([ops.sumsq_partials(x) for x in xs])
        out5 = torch.empty(5, dtype=torch.float32, device=xs[0].device)
        ops.loss_assemble(None, 0, 0, 0.0, parts, None, 0.5 * coef, None, None, 0, 0.0, out5)
        ctx.coef = coef
        ctx.save_for_backward(*xs)
        return out5[3]

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().float()
        outs = []
        for x in ctx.saved_tensors:
            o = torch.empty(x.shape, dtype=torch.float32, device=x.device)
            ops.axpby(x, ctx.coef, 0.0, o, alpha_dev=g)
            outs.append(o)
        return (None, *outs)


def feat_reg_loss(g_item_image, g_item_text, g_user_image, g_user_text, n_items: int, feat_reg_decay: float = 1e-5):
    """Trainer.feat_reg_loss_calculation (main.py:252-257): decay * 1/2 * sum of squares / n_items."""
    return _FeatRegFn.apply(float(feat_reg_decay) / float(n_items), g_item_image, g_item_text, g_user_image, g_user_text)
