"""Serving collectives fused into the shard kernels (world > 1, CUDA).

The reference answers ``pull`` / ``pullAverage`` / ``norms`` / ``multiply`` by
fanning a request out to the S parameter servers and concatenating or summing
the S replies on the client (`BigWord2VecMatrix` [G]; call sites MLLIB:486,514,
598, ML:353,453); ``findSynonyms`` then scans the V scores on the driver
(MLLIB:600-617).  Here every shard kernel stores its result directly into its
peers' symmetric memory over NVLink and publishes a sequence number
(``csrc/serve_fused.cu``, ``csrc/serve_common.cuh``); the score GEMM's epilogue
is a reduce-scatter (row ``v`` belongs to rank ``v // vown``), each owner
ranks only its ``V/S`` rows and the winners are exchanged.  No NCCL call.

Buffer discipline: every operation is ``push -> wait -> consume -> barrier``
on the current stream, so a peer can only overwrite a buffer after this rank
has consumed it.  All ranks must issue the same serving calls in the same
order (SPMD), which the shard-server request broadcast guarantees.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import torch

from . import cuda as _cuda_mod


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class ServeExchange:
    """Symmetric workspace + sequence context of one engine's serving path (collective constructor)."""

    ROW_CAP = 16384                 # rows per pull / sentences per pullAverage push
    CAND_CAP = 1 << 22              # candidate entries (world * Q * k) per exchange

    def __init__(self, ops):
        from ..parallel.symm import alloc_symmetric
        self.ops = ops
        self.C = _cuda_mod.extension()
        e = ops.e
        self.world, self.rank = ops.world, ops.rank
        self.K = ops.K
        self.V = int(ops.cfg.vocab_size)
        self.ldo = self.world * self.K
        self.vown = _round_up((self.V + self.world - 1) // self.world, 128)
        self.nvalid = max(0, min(self.V, (self.rank + 1) * self.vown) - self.rank * self.vown)
        vpad = self.vown * self.world
        slab_mb = int(os.environ.get("GW2V_SERVE_SLAB_MB", "1024"))
        self.q_cap = int(max(1, min(256, (slab_mb << 20) // (vpad * 4))))
        # region layout (bytes, 256-aligned)
        off = 0
        self.off = {}

        def region(name, nbytes):
            nonlocal off
            self.off[name] = off
            off += _round_up(nbytes, 256)

        region("flags", 256)
        region("rows", self.ROW_CAP * self.ldo * 4)
        region("slab", self.q_cap * vpad * 4)
        region("norms", vpad * 4)
        region("vec", vpad * 4)
        region("cand_v", self.CAND_CAP * 4)
        region("cand_i", self.CAND_CAP * 8)
        self.buf = alloc_symmetric(off, ops.dev, e.comm.group)
        self.local = self.buf.local
        self.done = torch.zeros(1, dtype=torch.int32, device=ops.dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=ops.dev)
        self.ctx = self.C.ServeCtx(self.world, self.rank, self.ptrs("flags"), self.done, self.err)
        self.have_norms = False

    # ------------------------------------------------------------------ helpers
    def ptrs(self, name: str) -> List[int]:
        return [p + self.off[name] for p in self.buf.ptrs]

    def view(self, name: str, numel: int, dtype=torch.float32) -> torch.Tensor:
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        return self.local[self.off[name]:self.off[name] + nbytes].view(dtype)

    # ------------------------------------------------------------------ K8 / K9
    def pull(self, rows_dev: torch.Tensor) -> torch.Tensor:
        """[R, world*K] full (padded) rows; the column all-gather happens inside the gather kernel."""
        syn0 = self.ops.e.syn0
        outs = []
        for lo in range(0, max(1, rows_dev.numel()), self.ROW_CAP):
            chunk = rows_dev[lo:lo + self.ROW_CAP].contiguous()
            self.C.serve_gather_push(self.ctx, syn0, chunk, self.ptrs("rows"), self.ldo)
            outs.append(self.view("rows", chunk.numel() * self.ldo).view(chunk.numel(), self.ldo).clone())
            self.ctx.barrier()
            self.ops.launches += 3
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    def pull_average(self, rows_flat: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
        syn0 = self.ops.e.syn0
        ns = offsets.numel() - 1
        outs = []
        for lo in range(0, max(1, ns), self.ROW_CAP):
            hi = min(ns, lo + self.ROW_CAP)
            offs = offsets[lo:hi + 1].contiguous()
            self.C.serve_segment_mean_push(self.ctx, syn0, rows_flat, offs, self.ptrs("rows"), self.ldo)
            outs.append(self.view("rows", (hi - lo) * self.ldo).view(hi - lo, self.ldo).clone())
            self.ctx.barrier()
            self.ops.launches += 3
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    # ------------------------------------------------------------------ K10
    def norms(self) -> torch.Tensor:
        """Euclidean row norms [V]: reduce-scatter of the partial sums of squares, sqrt on the owner,
        all-gather of the owner slices - three kernels, all traffic in-kernel."""
        syn0 = self.ops.e.syn0
        self.C.serve_sqnorm_push(self.ctx, syn0, self.ptrs("slab"), self.vown)
        slab = self.view("slab", self.world * self.vown)
        self.C.serve_reduce_finish_push(self.ctx, slab, self.world, self.vown, self.nvalid, True, self.ptrs("norms"))
        self.ctx.barrier()
        self.ops.launches += 5
        self.have_norms = True
        return self.view("norms", self.world * self.vown)[:self.V]

    def norms_owned(self) -> torch.Tensor:
        if not self.have_norms:
            self.norms()
        return self.view("norms", self.world * self.vown)[self.rank * self.vown:(self.rank + 1) * self.vown]

    def invalidate(self):
        self.have_norms = False

    # ------------------------------------------------------------------ K11
    def multiply(self, qs: torch.Tensor) -> torch.Tensor:
        """``syn0 @ q`` for ONE query slice [1, K] -> [V] (exact fp32)."""
        syn0 = self.ops.e.syn0
        self.C.serve_scores_push(self.ctx, syn0, qs.contiguous(), self.ptrs("slab"), self.vown, False)
        slab = self.view("slab", self.world * self.vown)
        self.C.serve_reduce_finish_push(self.ctx, slab, self.world, self.vown, self.nvalid, False, self.ptrs("vec"))
        out = self.view("vec", self.world * self.vown)[:self.V].clone()
        self.ctx.barrier()
        self.ops.launches += 5
        return out

    def top_k(self, qs: torch.Tensor, k: int, use_tc: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        """Cosine top-k of the query slices [Q, K] -> (idx [Q, k], sim [Q, k]); identical on every rank."""
        nq = qs.shape[0]
        k = int(k)
        per_q = self.world * k
        if per_q > self.CAND_CAP:
            raise ValueError(f"top-k of {k} is beyond the candidate exchange capacity")
        q_step = max(1, min(self.q_cap, self.CAND_CAP // per_q))
        norms_owned = self.norms_owned()
        syn0 = self.ops.e.syn0
        idx_out, sim_out = [], []
        for lo in range(0, nq, q_step):
            q = qs[lo:lo + q_step].contiguous()
            n = q.shape[0]
            tc = use_tc and bool(self.C.scores_tc_supported(self.K, n))
            self.C.serve_scores_push(self.ctx, syn0, q, self.ptrs("slab"), self.vown, tc)
            slab = self.view("slab", self.world * n * self.vown)
            self.C.serve_topk_owned_push(self.ctx, slab, self.world, n, self.vown, self.nvalid, norms_owned,
                                         self.rank * self.vown, k, self.ptrs("cand_v"), self.ptrs("cand_i"))
            cv = self.view("cand_v", n * per_q).view(n, per_q)
            ci = self.view("cand_i", n * per_q, torch.int64).view(n, per_q)
            idx, sim = self.C.serve_topk_final(cv, ci, k)
            self.ctx.barrier()
            self.ops.launches += 8
            idx_out.append(idx)
            sim_out.append(sim)
        if len(idx_out) == 1:
            return idx_out[0], sim_out[0]
        return torch.cat(idx_out, 0), torch.cat(sim_out, 0)

    def allgather_sum(self, part: torch.Tensor) -> torch.Tensor:
        """Sum of a small fp32 tensor over ranks in fixed rank order (identical bits on every rank)."""
        flat = part.contiguous().view(-1)
        n = flat.numel()
        if n * self.world > self.CAND_CAP:
            raise ValueError("allgather_sum payload too large")
        self.C.serve_push_block(self.ctx, flat, self.ptrs("cand_v"))
        out = self.view("cand_v", self.world * n).view(self.world, n).sum(0).view(part.shape)
        self.ctx.barrier()
        self.ops.launches += 3
        return out
