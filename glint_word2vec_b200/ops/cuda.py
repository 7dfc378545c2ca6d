"""CUDA (sm_100a) execution of the shard-engine operations.

Everything here dispatches to the hand-written kernels in ``csrc/`` through the
in-tree extension ``glint_word2vec_b200/_C.so``.  There is deliberately NO
PyTorch fallback on a GPU: if the extension is missing we fail loudly so a
silent eager path can never masquerade as the product.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ..utils import philox

try:
    from .. import _C                      # built by glint_word2vec_b200/build_ext.py
except ImportError as e:                  # pragma: no cover
    raise ImportError(
        "glint_word2vec_b200._C (the sm_100a kernel extension) is not built; run "
        "`python -m glint_word2vec_b200.build_ext` (or __graft_entry__.build()). "
        f"Original error: {e}") from e

U32_MAX = 0xFFFFFFFF
WINDOW_MODES = {"reference": 0, "word2vec_c": 1}


def extension():
    return _C


class CudaShardOps:
    """GPU state + kernels of one ``ShardEngine``."""

    def __init__(self, engine):
        self.e = engine
        self.dev = engine.device
        torch.cuda.set_device(self.dev)
        self.cfg = engine.cfg
        self.K = engine.shard.cols
        self.world = engine.comm.world
        self.rank = engine.comm.rank
        self.alias_dev: Optional[torch.Tensor] = None
        self.keep_dev: Optional[torch.Tensor] = None
        self.subsample_active = True
        self._cap = 0
        self._epoch = 0
        self._pg_epoch = 0
        self._stats_ring = None
        self._stats_i = 0
        self._xchg = None
        self.compute_loss = True
        self.timing: Optional[torch.Tensor] = None
        self.launches = 0                    # kernel launches issued by this object (bench bookkeeping)
        self._count_val = -1
        self.debug = int(os.environ.get("GW2V_DEBUG", "0"))      # profiling-only kernel switches
        # sigmoid_mode="table": the reference's 1000-entry sigma table (MLLIB:281-302), looked up in-kernel
        self.exp_table = None
        if engine.cfg.sigmoid_mode == "table":
            from ..models.sgns import _exp_table
            self.exp_table = _exp_table().to(self.dev).contiguous()
        self._props = torch.cuda.get_device_properties(self.dev)
        # world > 1: serving collectives are fused into the kernels (ops/serving.py); GW2V_SERVE_FUSED=0 falls
        # back to kernel + NCCL collective (kept for A/B measurements)
        self.serve_fused = self.world > 1 and os.environ.get("GW2V_SERVE_FUSED", "1") != "0"
        self._serve = None

    # ------------------------------------------------------------------ setup
    def init_weights(self, seed: int) -> Tuple[torch.Tensor, torch.Tensor]:
        v = self.cfg.vocab_size
        syn0 = torch.empty(v, self.K, dtype=torch.float32, device=self.dev)
        _C.init_syn0(syn0, self.rank * self.K, self.cfg.vector_size, int(seed))
        self.launches += 1
        syn1 = torch.zeros(v, self.K, dtype=torch.float32, device=self.dev)
        return syn0, syn1

    def upload_noise(self, alias, keep_thresh: np.ndarray):
        self.alias_dev = torch.from_numpy(alias.packed()).to(self.dev)
        self.keep_dev = torch.from_numpy(keep_thresh.view(np.int32).copy()).to(self.dev)
        self.subsample_active = bool((keep_thresh != U32_MAX).any())

    def release(self):
        self._xchg = None
        self._serve = None
        self.alias_dev = self.keep_dev = None
        self._cap = 0

    def _ensure_capacity(self, t: int):
        if t <= self._cap:
            return
        cap = max(t, 1024)
        d = self.dev
        self.tok_in = torch.empty(cap, dtype=torch.int32, device=d)
        self.sid_in = torch.empty(cap, dtype=torch.int32, device=d)
        self.tok_c = torch.empty(cap, dtype=torch.int32, device=d)
        self.sid_c = torch.empty(cap, dtype=torch.int32, device=d)
        self.count = torch.zeros(1, dtype=torch.int32, device=d)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=d)
        self.chain = torch.zeros(int(_C.subsample_max_blocks(cap)) + 1, dtype=torch.int64, device=d)
        self.pin_tok = torch.empty(cap, dtype=torch.int32).pin_memory()
        self.pin_sid = torch.empty(cap, dtype=torch.int32).pin_memory()
        self._stats_ring = torch.zeros(256, 4, dtype=torch.float32, device=d)
        self._count_val = -1
        # pair-generation workspaces (pairgen.cu)
        cfg = self.cfg
        self.pd = int(_C.pairgen_desc_ints(cfg.negatives))
        self.pg_cinfo = torch.empty(cap, dtype=torch.int32, device=d)
        self.pg_off = torch.empty(cap, dtype=torch.int32, device=d)
        self.pg_npairs = torch.zeros(1, dtype=torch.int32, device=d)
        self.pg_desc = torch.empty(cap * 2 * cfg.window * self.pd, dtype=torch.int32, device=d)
        self.pg_ticket = torch.zeros(1, dtype=torch.int32, device=d)
        self.pg_chain = torch.zeros(int(_C.pairgen_max_blocks(cap)) + 1, dtype=torch.int64, device=d)
        self._cap = cap

    # ------------------------------------------------------------------ cross-shard exchange
    def _setup_exchange(self):
        """Symmetric exchange slots + flags for the in-kernel all-reduce (collective)."""
        from ..parallel.symm import alloc_symmetric
        cfg = self.cfg
        dev_index = self.dev.index or 0
        want = os.environ.get("GW2V_MULTI_KERNEL", "auto")
        pipe_ok = bool(_C.sgns_pipe_multi_supported(self.K, cfg.window, cfg.negatives))
        group_ok = bool(_C.sgns_group_multi_supported(self.K, cfg.window, cfg.negatives))
        pairs_ok = bool(_C.sgns_pairs_supported(self.K, cfg.window, cfg.negatives))
        if want == "auto":
            want = "pairs" if pairs_ok else ("group" if group_ok else ("pipe" if pipe_ok else "v1"))
        variant = 3 if (want == "pairs" and pairs_ok) else (
            2 if (want == "group" and group_ok) else (1 if (want == "pipe" and pipe_ok) else 0))
        if variant == 3:
            grid = int(_C.sgns_pairs_grid(self.K, dev_index, True))
            warps, nslot, slot_floats = [int(x) for x in _C.sgns_pairs_multi_geometry()]
            tb = 0
            units = grid * warps
            xbytes = units * nslot * self.world * slot_floats * 4
            nseq = units
        elif variant >= 1:
            geo = _C.sgns_group_multi_geometry(self.K, dev_index) if variant == 2 else \
                _C.sgns_pipe_multi_geometry(self.K, cfg.negatives, dev_index)
            grid, warps, nslot, slot_floats = [int(x) for x in geo]
            tb = 0
            units = grid * warps                      # one exchange ring per warp
            xbytes = units * nslot * self.world * slot_floats * 4
            nseq = units
        else:
            tb = int(os.environ.get("GW2V_TILE_CENTERS", "16"))
            grid = int(_C.sgns_multi_max_grid(self.K, cfg.window, cfg.negatives, tb, dev_index))
            maxpairs = tb * 2 * cfg.window
            slot_floats = (maxpairs * (1 + cfg.negatives) + 3) // 4 * 4
            units = grid                              # one exchange ring (2 slots) per CTA
            xbytes = units * 2 * self.world * slot_floats * 4
            nseq = grid
        # all ranks must launch the identical geometry (the flag protocol pairs warp w with warp w)
        g = torch.tensor([grid, -grid], dtype=torch.int64, device=self.dev)
        dist.all_reduce(g, op=dist.ReduceOp.MIN, group=self.e.comm.group)
        if int(g[0].item()) != grid or int(-g[1].item()) != grid:
            raise RuntimeError("ranks disagree on the persistent grid size; heterogeneous GPUs are not supported")
        fbytes = (units * self.world * 4 + 255) // 256 * 256
        xbytes = (xbytes + 255) // 256 * 256
        buf = alloc_symmetric(xbytes + fbytes, self.dev, self.e.comm.group)
        self._xchg = {
            "buf": buf, "grid": grid, "tb": tb, "slot_floats": slot_floats, "variant": variant,
            "xptrs": list(buf.ptrs), "fptrs": [p + xbytes for p in buf.ptrs],
            "mc": buf.multicast_ptr,
            # NVLS multicast push (multimem.st) is opt-in: GW2V_NVLS=1 and a multicast mapping granted by the driver
            "mc_x": buf.multicast_ptr if (buf.multicast_ptr and os.environ.get("GW2V_NVLS", "0") == "1") else 0,
            "mc_f": (buf.multicast_ptr + xbytes) if (buf.multicast_ptr and os.environ.get("GW2V_NVLS", "0") == "1") else 0,
            "cta_seq": torch.zeros(nseq, dtype=torch.int32, device=self.dev),
            "err": torch.zeros(1, dtype=torch.int32, device=self.dev),
        }
        self.timing = torch.zeros(2, dtype=torch.int64, device=self.dev)

    # ------------------------------------------------------------------ training
    def stage_tokens(self, tokens, sent_id) -> int:
        """Host -> device copy of one step's inputs through pinned memory."""
        t = int(len(tokens))
        self._ensure_capacity(t)
        if isinstance(tokens, torch.Tensor) and tokens.is_pinned():
            self.tok_in[:t].copy_(tokens, non_blocking=True)
            self.sid_in[:t].copy_(sent_id, non_blocking=True)
        else:
            self.pin_tok[:t].copy_(torch.as_tensor(np.ascontiguousarray(tokens, dtype=np.int32)))
            self.pin_sid[:t].copy_(torch.as_tensor(np.ascontiguousarray(sent_id, dtype=np.int32)))
            self.tok_in[:t].copy_(self.pin_tok[:t], non_blocking=True)
            self.sid_in[:t].copy_(self.pin_sid[:t], non_blocking=True)
        return t

    def train_step(self, tokens, sent_id, raw_pos0: int, iteration: int, alpha: float) -> torch.Tensor:
        t = self.stage_tokens(tokens, sent_id)
        return self.train_step_staged(t, raw_pos0, iteration, alpha)

    def train_step_staged(self, t: int, raw_pos0: int, iteration: int, alpha: float) -> torch.Tensor:
        return self.train_step_device(self.tok_in, self.sid_in, t, raw_pos0, iteration, alpha)

    def train_step_device(self, tok_dev: torch.Tensor, sid_dev: torch.Tensor, t: int, raw_pos0: int,
                          iteration: int, alpha: float) -> torch.Tensor:
        """One step on tokens that already live on the device (int32 tensors of length >= t)."""
        cfg = self.cfg
        e = self.e
        self._ensure_capacity(t)
        if self.world > 1 and self._xchg is None:
            self._setup_exchange()
        if self.subsample_active:
            self._epoch += 1
            _C.subsample_compact(tok_dev, sid_dev, t, self.keep_dev, int(cfg.seed), int(iteration),
                                 int(raw_pos0), self.tok_c, self.sid_c, self.count, self.ticket, self.chain,
                                 self._epoch)
            self.launches += 3            # count, tile scan, scatter
            self._count_val = -1
            tok, sid = self.tok_c, self.sid_c
        else:
            if self._count_val != t:
                self.count.fill_(t)
                self._count_val = t
            tok, sid = tok_dev, sid_dev
        self._stats_i = (self._stats_i + 1) % self._stats_ring.shape[0]
        stats = self._stats_ring[self._stats_i]
        stats.zero_()
        wm = WINDOW_MODES[cfg.window_mode]
        if self.world > 1 and self._xchg["variant"] == 3:
            x = self._xchg
            self._pg_epoch += 1
            _C.sgns_step_pairs(e.syn0, e.syn1, tok, sid, self.count, t, self.alias_dev, stats, int(raw_pos0),
                               int(cfg.seed), int(iteration), cfg.window, cfg.negatives, wm, float(alpha),
                               float(cfg.max_grad), self.compute_loss, x["grid"], self.world, self.rank,
                               x["xptrs"], x["fptrs"], x["cta_seq"], x["err"], self.timing, self.debug,
                               self.pg_cinfo, self.pg_off, self.pg_npairs, self.pg_desc, self.pg_ticket,
                               self.pg_chain, self._pg_epoch, x["mc_x"], x["mc_f"], self.exp_table)
            self.launches += 3            # + the training kernel counted below
        elif self.world > 1:
            x = self._xchg
            _C.sgns_step(e.syn0, e.syn1, tok, sid, self.count, self.alias_dev, stats, int(raw_pos0),
                         int(cfg.seed), int(iteration), cfg.window, cfg.negatives, wm, float(alpha),
                         float(cfg.max_grad), self.compute_loss, x["grid"], self.world, self.rank, x["tb"],
                         x["slot_floats"], x["xptrs"], x["fptrs"], x["mc"], x["cta_seq"], x["err"], self.timing,
                         self.debug, x["variant"], self.exp_table)
        else:
            if not hasattr(self, "_grid1"):
                self._variant, self._grid1 = self._pick_single_kernel()
            if self._variant == 3:
                self._pg_epoch += 1
                _C.sgns_step_pairs(e.syn0, e.syn1, tok, sid, self.count, t, self.alias_dev, stats, int(raw_pos0),
                                   int(cfg.seed), int(iteration), cfg.window, cfg.negatives, wm, float(alpha),
                                   float(cfg.max_grad), self.compute_loss, self._grid1, 1, 0, [], [], None, None,
                                   None, self.debug, self.pg_cinfo, self.pg_off, self.pg_npairs, self.pg_desc,
                                   self.pg_ticket, self.pg_chain, self._pg_epoch, 0, 0, self.exp_table)
                self.launches += 4            # pair_count, pair_tile_scan, pair_fill, sgns_pairs
                return stats
            _C.sgns_step(e.syn0, e.syn1, tok, sid, self.count, self.alias_dev, stats, int(raw_pos0),
                         int(cfg.seed), int(iteration), cfg.window, cfg.negatives, wm, float(alpha),
                         float(cfg.max_grad), self.compute_loss, self._grid1, 1, 0, 0, 0, [], [], 0,
                         None, None, None, self.debug, self._variant, self.exp_table)
        self.launches += 1
        return stats

    def _top_k_fused(self, qs: torch.Tensor, norms: torch.Tensor, k: int, tc: bool):
        """world > 1: score GEMM with a reduce-scatter epilogue, per-owner top-k, candidate exchange -
        all inside the kernels.  With tf32 screening the k+16 survivors are re-scored in exact fp32."""
        sx = self.serve()
        v = self.cfg.vocab_size
        nq = qs.shape[0]
        if not tc:
            return sx.top_k(qs, k, False)
        kk = min(v, k + 16)
        idx, _ = sx.top_k(qs, kk, True)
        safe = idx.clamp(min=0)
        rows = _C.gather_rows(self.e.syn0, safe.reshape(-1).contiguous()).view(nq, kk, self.K)
        self.launches += 1
        dots = sx.allgather_sum((rows * qs[:, None, :]).sum(-1))          # exact fp32, fixed rank order
        nr = norms[safe]
        cos = torch.where((nr > 0) & (idx >= 0), dots / nr.clamp(min=1e-30), torch.zeros_like(dots))
        cos = torch.where(idx >= 0, cos, torch.full_like(cos, -3.0e38))
        sim, order = torch.sort(cos, dim=1, descending=True)
        return torch.gather(idx, 1, order)[:, :k].contiguous(), sim[:, :k].contiguous()

    def _pick_single_kernel(self):
        """single-shard kernel variant: 2 = lane-group register path, 1 = TMA pipeline, 0 = v1."""
        cfg = self.cfg
        dev_index = self.dev.index or 0
        want = os.environ.get("GW2V_SINGLE_KERNEL", "auto")
        group_ok = bool(_C.sgns_group_supported(self.K, cfg.window, cfg.negatives))
        pipe_ok = bool(_C.sgns_pipe_supported(self.K, cfg.window, cfg.negatives))
        pairs_ok = bool(_C.sgns_pairs_supported(self.K, cfg.window, cfg.negatives))
        if want == "auto":
            # measured on B200 (profiles/r1_kernel_variants.md): the lane-group register path wins at every
            # row length (TMA bulk copies cost ~25 SM cycles each, which binds the pipeline for short rows)
            want = "pairs" if pairs_ok else ("group" if group_ok else ("pipe" if pipe_ok else "v1"))
        if want == "pairs" and pairs_ok:
            return 3, int(_C.sgns_pairs_grid(self.K, dev_index, False))
        if want == "group" and group_ok:
            return 2, int(_C.sgns_group_grid(self.K, dev_index))
        if want == "pipe" and pipe_ok:
            return 1, int(_C.sgns_pipe_grid(self.K, cfg.negatives, dev_index))
        return 0, int(_C.sgns_single_grid(self.K, dev_index))

    # ------------------------------------------------------------------ inference
    def serve(self):
        """Lazily allocated symmetric serving workspace (collective on first use)."""
        if self._serve is None:
            from .serving import ServeExchange
            self._serve = ServeExchange(self)
        return self._serve

    def _rows_dev(self, rows: torch.Tensor) -> torch.Tensor:
        return rows.to(self.dev, torch.int64).contiguous()

    def gather_rows(self, rows: torch.Tensor) -> torch.Tensor:
        self.launches += 1
        return _C.gather_rows(self.e.syn0, self._rows_dev(rows))

    def segment_mean_rows(self, rows_flat: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
        self.launches += 1
        return _C.segment_mean_rows(self.e.syn0, self._rows_dev(rows_flat), self._rows_dev(offsets))

    def row_sqnorm(self) -> torch.Tensor:
        self.launches += 1
        return _C.row_sqnorm(self.e.syn0)

    def _use_tc(self, nq: int) -> bool:
        """tcgen05 (tf32) screening pays off once the CUDA-core sweep turns compute bound (Q >= 8)."""
        mode = os.environ.get("GW2V_NN_TC", "auto")
        if mode == "0" or not _C.scores_tc_supported(self.K, nq):
            return False
        return mode == "1" or nq >= int(os.environ.get("GW2V_NN_TC_MIN_Q", "8"))

    def scores(self, qs: torch.Tensor, allow_tc: bool = False) -> torch.Tensor:
        """Partial scores [Q, V] of this shard's columns (exact fp32 unless ``allow_tc``)."""
        qs = qs.contiguous()
        self.launches += 1
        if allow_tc and self._use_tc(qs.shape[0]):
            return _C.scores_tc(self.e.syn0, qs)
        outs = []
        maxq = max(1, (192 * 1024) // (self.K * 4))
        for lo in range(0, qs.shape[0], maxq):
            outs.append(_C.scores_rows(self.e.syn0, qs[lo:lo + maxq].contiguous()))
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    def top_k(self, qs: torch.Tensor, norms: torch.Tensor, k: int):
        """Cosine top-k.  Large query batches are screened on the tensor cores (tf32) with a
        margin, then the few surviving candidates are re-scored in exact fp32, so reported
        similarities never carry tf32 rounding."""
        nq = qs.shape[0]
        v = self.cfg.vocab_size
        tc = self._use_tc(nq)
        if self.serve_fused:
            return self._top_k_fused(qs.contiguous(), norms, int(k), tc)
        part = self.scores(qs, allow_tc=tc)
        full = self.e.comm.all_reduce_sum(part)
        self.launches += 1
        if not tc:
            return _C.cosine_topk(full, norms.contiguous(), int(k))
        kk = min(v, int(k) + 16)
        idx, _ = _C.cosine_topk(full, norms.contiguous(), kk)            # [Q, kk] candidates
        rows = _C.gather_rows(self.e.syn0, idx.reshape(-1).contiguous()).view(nq, kk, self.K)
        self.launches += 1
        dots = (rows * qs[:, None, :]).sum(-1)                           # exact fp32 partial dots (tiny)
        dots = self.e.comm.all_reduce_sum(dots.contiguous())
        nr = norms[idx]
        cos = torch.where(nr > 0, dots / nr.clamp(min=1e-30), torch.zeros_like(dots))
        sim, order = torch.sort(cos, dim=1, descending=True)
        return torch.gather(idx, 1, order)[:, :k].contiguous(), sim[:, :k].contiguous()
