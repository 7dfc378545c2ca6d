"""The number to beat: the same column-sharded SGNS step written with stock
torch ops + ``torch.distributed.all_reduce`` (NCCL).

The reference itself cannot be built here (Scala/sbt/Spark/Glint, no JVM, no
network, no GPU code at all -- BASELINE.md section 2), so per BASELINE.json the
comparison target is "an NCCL(+stock ops) build of the same column-sharded
run": identical sharding, identical pairs/negatives (same Philox streams), with
``index_select`` -> row.row dot -> ``all_reduce`` -> sigmoid -> ``index_add_``.
This is the baseline, never the product.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from glint_word2vec_b200.models import sgns


class BaselineShard:
    """Unfused column shard: Glint-style ``dotprod`` / ``adjust`` with library ops."""

    def __init__(self, engine):
        self.e = engine                       # reuses the engine's weights + noise tables

    def step(self, tokens: np.ndarray, sent_id: np.ndarray, pos0: int, iteration: int, alpha: float,
             pairs_host=None):
        e = self.e
        cfg = e.cfg
        dev = e.device
        if pairs_host is None:
            pairs_host = self.enumerate(tokens, sent_id, pos0, iteration)
        w, c, ng = (x.to(dev, non_blocking=True) for x in pairs_host)
        u = e.syn0.index_select(0, w)
        vc = e.syn1.index_select(0, c)
        vn = e.syn1.index_select(0, ng.reshape(-1)).view(ng.shape[0], ng.shape[1], -1)
        f = torch.empty(w.shape[0], 1 + ng.shape[1], device=dev)
        f[:, 0] = (u * vc).sum(-1)
        f[:, 1:] = torch.einsum("pd,pnd->pn", u, vn)
        if e.comm.world > 1:
            dist.all_reduce(f, group=e.comm.group)
        mask = (ng != c[:, None]).float()
        gp = sgns.sigmoid_coeff(f[:, 0], 1.0, alpha)
        gm = sgns.sigmoid_coeff(f[:, 1:], 0.0, alpha) * mask
        du = gp[:, None] * vc + torch.einsum("pn,pnd->pd", gm, vn)
        e.syn1.index_add_(0, c, gp[:, None] * u)
        e.syn1.index_add_(0, ng.reshape(-1), (gm[:, :, None] * u[:, None, :]).reshape(-1, u.shape[1]))
        e.syn0.index_add_(0, w, du)
        return int(w.shape[0])

    def enumerate(self, tokens, sent_id, pos0, iteration):
        """Pair/negative enumeration on the host (not timed by the benchmark)."""
        cfg = self.e.cfg
        ci, cj, slot = sgns.enumerate_pairs(cfg, tokens, sent_id, pos0, iteration)
        pos = np.uint64(pos0) + ci.astype(np.uint64)
        neg = sgns.draw_negatives(cfg, self.e.alias, pos, slot, iteration)
        tok = tokens.astype(np.int64)
        return (torch.from_numpy(tok[ci]).pin_memory(), torch.from_numpy(tok[cj]).pin_memory(),
                torch.from_numpy(neg.astype(np.int64)).pin_memory())
