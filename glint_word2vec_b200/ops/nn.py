"""Nearest-neighbour index of a trained shard engine: row-sharded serving replica + fused select kernel.

Reference: ``findSynonyms`` = ``multiply`` on every parameter server (one sgemv over its column slice), a client-side
sum of the S partial score vectors, division by the norms and a driver-side scan of all V cosines (MLLIB:589-617).
Column shards are the right layout for training (every pair touches every shard, dots are all-reduced in-kernel) but
the wrong one for search: every shard streams all V rows and the V x Q partial scores have to cross NVLink.

For search each rank therefore keeps a ROW shard of the input vectors: rows ``[rank * vown, (rank + 1) * vown)`` with
all columns, built from the column shards by one peer-store kernel (``rowshard_push_kernel``) whenever the weights
have changed -- the same number of bytes per GPU as its column shard.  A query batch then costs each GPU one sweep of
``V / S`` full rows (the fused tcgen05 select kernel, ``csrc/nn_select.cu``), an exact fp32 re-score of the ~k*stride
survivors per query, a local top-k, and an exchange of ``k`` (value, index) pairs per query and rank.  Nothing of
size V ever leaves a GPU, and S GPUs stream S times the bytes per second.

Single GPU: the index is syn0 itself when the row length is a multiple of 32 floats, otherwise a zero-padded copy.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import cuda as _cuda_mod

#: cosine error budget of the tf32 screening pass (|a.b - tf32(a).tf32(b)| <= 2^-10 for unit vectors), applied twice
#: (threshold sample + sweep) with a safety factor
TF32_MARGIN = 4.0e-3
CAND_CAP = 4096          # candidate slots per query
MIN_SELECT_ROWS = 1 << 17


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class NNIndex:
    def __init__(self, ops):
        self.ops = ops
        self.C = _cuda_mod.extension()
        self.version = -1
        self.mat: Optional[torch.Tensor] = None          # [rows, Kp]
        self.inv: Optional[torch.Tensor] = None          # [rows] 1/|row| (0 for zero rows)
        self.row_base = 0
        self.rows = 0
        self._symm = None
        self.overflows = 0                               # batches that fell back to the dense path
        e = ops.e
        self.world, self.rank = ops.world, ops.rank
        self.V = int(ops.cfg.vocab_size)
        self.d = int(ops.cfg.vector_size)
        self.K = ops.K
        self.Kp = _round_up(self.world * self.K, 32)
        if e.comm.world > 1:
            self.vown = ops.serve().vown
            self.row_base = self.rank * self.vown
            self.rows = max(0, min(self.V, (self.rank + 1) * self.vown) - self.row_base)
        else:
            self.vown = self.V
            self.rows = self.V

    # ------------------------------------------------------------------ build
    def refresh(self):
        """(Re)build the replica if the weights changed since the last build (collective when world > 1)."""
        e = self.ops.e
        if self.version == e._version and self.mat is not None:
            return
        syn0 = e.syn0
        if e.comm.world > 1:
            sx = self.ops.serve()
            if self._symm is None:
                from ..parallel.symm import alloc_symmetric
                self._symm = alloc_symmetric(self.vown * self.Kp * 4, self.ops.dev, e.comm.group)
                self._symm.local.zero_()                 # padding columns / rows beyond V stay zero for good
                torch.cuda.current_stream(self.ops.dev).synchronize()
            sx.ctx.barrier()                             # nobody still reads the previous replica
            self.C.serve_rowshard_push(sx.ctx, syn0, list(self._symm.ptrs), self.vown, self.Kp)
            sx.ctx.barrier()
            self.ops.launches += 4
            self.mat = self._symm.local.view(torch.float32).view(self.vown, self.Kp)[:max(self.rows, 1)]
        elif self.K % 32 == 0:
            self.mat = syn0                              # live view, no copy
        else:
            if self.mat is None or self.mat.data_ptr() == syn0.data_ptr():
                self.mat = torch.zeros(self.V, self.Kp, dtype=torch.float32, device=self.ops.dev)
            self.mat[:, :self.K].copy_(syn0)
        sq = self.C.row_sqnorm(self.mat)
        self.ops.launches += 1
        self.inv = torch.where(sq > 0, torch.rsqrt(sq.clamp(min=1e-37)), torch.zeros_like(sq)).contiguous()
        self.version = e._version

    def release(self):
        self.mat = self.inv = self._symm = None
        self.version = -1

    # ------------------------------------------------------------------ query
    def supported(self, nq: int, k: int) -> bool:
        if os.environ.get("GW2V_NN_SELECT", "1") == "0":
            return False
        return (self.V >= MIN_SELECT_ROWS * self.world and k * 4 <= CAND_CAP // 4
                and bool(self.C.nn_select_supported(self.Kp, min(nq, 256))))

    def _local_candidates(self, q: torch.Tensor, k: int):
        """q [n <= 256, Kp] unit queries -> exact candidates (values [n, CAND_CAP], global indices [n, CAND_CAP])."""
        n = q.shape[0]
        qpad = self.C.nn_pad_queries(q)
        rows = self.rows
        mat = self.mat[:rows]
        # threshold: k-th best cosine among every `stride`-th row (a subset, hence a lower bound of the true k-th best)
        stride = int(max(1, min(64, CAND_CAP // (4 * k), rows // 65536)))
        kk = min(k, (rows + stride - 1) // stride)
        sample = self.C.nn_sample_cosines(mat, self.inv, qpad, n, stride)
        thr = torch.topk(sample, kk, dim=1).values[:, -1].contiguous() - TF32_MARGIN
        if kk < k:
            thr.fill_(-3.0e38)
        cand, count = self.C.nn_select(mat, self.inv, qpad, n, thr, CAND_CAP)
        val, idx = self.C.nn_rerank(mat, self.inv, qpad, n, cand, count, self.row_base)
        self.ops.launches += 3
        return val, idx, count

    def top_k(self, q_unit: torch.Tensor, k: int) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
        """Cosine top-k of unit-length queries [Q, d] (identical on every rank).  Returns device tensors
        (idx [Q, k], sim [Q, k]), or None if a candidate list overflowed (caller falls back to the dense path)."""
        self.refresh()
        dev = self.ops.dev
        nq = q_unit.shape[0]
        q = torch.zeros(nq, self.Kp, dtype=torch.float32, device=dev)
        q[:, :self.d] = q_unit.to(dev)
        idx_out, sim_out, counts = [], [], []
        multi = self.ops.e.comm.world > 1
        sx = self.ops.serve() if multi else None
        q_step = 256 if not multi else max(1, min(256, sx.CAND_CAP // (self.world * k)))
        for lo in range(0, nq, q_step):
            qq = q[lo:lo + q_step].contiguous()
            n = qq.shape[0]
            if self.rows > 0:
                val, idx, count = self._local_candidates(qq, k)
            else:                                         # a rank without rows still takes part in the exchange
                val = torch.full((n, 32), -3.0e38, device=dev)
                idx = torch.full((n, 32), -1, dtype=torch.int64, device=dev)
                count = torch.zeros(n, dtype=torch.int32, device=dev)
            counts.append(count)
            if multi:
                self.C.serve_topk_cand_push(sx.ctx, val, idx, k, sx.ptrs("cand_v"), sx.ptrs("cand_i"))
                per_q = self.world * k
                cv = sx.view("cand_v", n * per_q).view(n, per_q)
                ci = sx.view("cand_i", n * per_q, torch.int64).view(n, per_q)
                i2, s2 = self.C.serve_topk_final(cv, ci, k)
                sx.ctx.barrier()
                self.ops.launches += 5
            else:
                i2, s2 = self.C.serve_topk_final(val, idx, k)
                self.ops.launches += 1
            idx_out.append(i2)
            sim_out.append(s2)
        over = torch.cat(counts).max() > CAND_CAP
        if multi:                                         # every rank must take the same branch
            flag = sx.allgather_sum(over.to(torch.float32).view(1))
            over = flag > 0
        if bool(over.item()):
            self.overflows += 1
            return None
        if len(idx_out) == 1:
            return idx_out[0], sim_out[0]
        return torch.cat(idx_out, 0), torch.cat(sim_out, 0)
