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


N_STAGE = 4                 # input staging slots (host may run this many steps ahead before blocking)
STATS_RING = 256            # in-flight asynchronous statistics read-backs


class _StageSlot:
    def __init__(self, cap: int, dev):
        self.pin_tok = torch.empty(cap, dtype=torch.int32).pin_memory()
        self.pin_sid = torch.empty(cap, dtype=torch.int32).pin_memory()
        # numpy views of the pinned buffers: a host-side torch copy_ of 0.5 MB enters an OpenMP region and was measured
        # at ~2 ms on a many-core box (4.7 of the 6 ms per step of the fit loop); a plain numpy memcpy is ~50 us
        self.pin_tok_np, self.pin_sid_np = self.pin_tok.numpy(), self.pin_sid.numpy()
        self.tok = torch.empty(cap, dtype=torch.int32, device=dev)
        self.sid = torch.empty(cap, dtype=torch.int32, device=dev)
        self.h2d_ev = torch.cuda.Event()
        self.free_ev = None
        self._free_ev = torch.cuda.Event()


class StepHandle:
    """Result of an asynchronous step: ``result()`` -> CPU tensor ``[pairs, loss, max|dot|, kept_tokens]``."""

    def __init__(self, ring, ev, j):
        self._ring, self._ev, self._j = ring, ev, j
        self._val = None

    def result(self) -> torch.Tensor:
        if self._val is None:
            self._ev.synchronize()
            self._val = self._ring[self._j].clone()
        return self._val


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
        # profiling only: run the column-shard kernel of a pretended world of S on ONE GPU (its pushes land in
        # its own buffer, see sgns_pairs.cu debug bit 3) so that ncu / quick sweeps do not need a multi-GPU box
        self._loopback = int(os.environ.get("GW2V_LOOPBACK_WORLD", "0")) if engine.comm.world == 1 else 0
        if self._loopback > 1:
            self.world = self._loopback
        self.alias_dev: Optional[torch.Tensor] = None
        self.keep_dev: Optional[torch.Tensor] = None
        self.subsample_active = True
        self._cap = 0
        self._stats_ring = None
        self._stats_i = 0
        self._xchg = None
        self.compute_loss = True
        self.timing: Optional[torch.Tensor] = None
        self.launches = 0                    # kernel launches issued by this object (bench bookkeeping)
        self._count_val = -1
        self.debug = int(os.environ.get("GW2V_DEBUG", "0"))      # profiling-only kernel switches
        if self._loopback > 1:
            self.debug |= 8
        self._share_centre = 1 if engine.cfg.neg_sharing == "centre" else 0
        # neg_sharing="tile": the tcgen05 kernel (csrc/sgns_tile.cu); there is no other device implementation of it
        self._tile_mode = engine.cfg.neg_sharing == "tile"
        if self._tile_mode and not _C.sgns_tile_supported(self.K, engine.cfg.window, engine.cfg.negatives,
                                                          engine.cfg.tile_centres, engine.cfg.tile_negatives):
            raise ValueError('neg_sharing="tile" on a GPU needs tile_centres=128, tile_negatives in {32, 64}, '
                             "window <= 11 and a multiple of 4 columns per shard")
        if not self._tile_mode and not _C.sgns_pairs_supported(self.K, engine.cfg.window, engine.cfg.negatives):
            # no silent fall-back to a slower kernel: the limits are compile-time constants of csrc/pairgen.cu /
            # csrc/sgns_pairs.cu.  transport="nccl" (un-fused torch path) runs any shape.
            raise ValueError(f"the fused GPU step supports window <= 11, negatives <= 21 and at most 1024 columns per "
                             f"shard (got window={engine.cfg.window}, negatives={engine.cfg.negatives}, "
                             f"columns/shard={self.K}); use more shards or parameterServerConfig transport='nccl'")
        self.row_scale0: Optional[torch.Tensor] = None     # per-row update scales of the first H rows (hot-row damping)
        self.row_scale1: Optional[torch.Tensor] = None
        self._scale_window = -1
        self._scale_cache = {}
        self.tile_dbg: Optional[torch.Tensor] = None       # tests: dot products of tile 0
        # sigmoid_mode="table": the reference's 1000-entry sigma table (MLLIB:281-302), looked up in-kernel
        self.exp_table = None
        if engine.cfg.sigmoid_mode == "table":
            from ..models.sgns import _exp_table
            self.exp_table = _exp_table().to(self.dev).contiguous()
        self._props = torch.cuda.get_device_properties(self.dev)
        self._copy_stream = None
        self._stage = []
        # optional cap on queued steps (0 = unlimited).  Measured on 4 GPUs: pacing the host does not reduce the
        # in-kernel waits (0.51-0.58 ms free-running vs 0.55-0.68 ms with a cap of 2), so the default is off.
        self._max_inflight = int(os.environ.get("GW2V_MAX_INFLIGHT", "0"))
        self._step_events = [None] * max(1, self._max_inflight)
        self._step_i = 0
        # world > 1: serving collectives are fused into the kernels (ops/serving.py); GW2V_SERVE_FUSED=0 falls
        # back to kernel + NCCL collective (kept for A/B measurements)
        self.serve_fused = engine.comm.world > 1 and os.environ.get("GW2V_SERVE_FUSED", "1") != "0"
        self._serve = None
        self._nn = None

    # ------------------------------------------------------------------ setup
    def init_weights(self, seed: int) -> Tuple[torch.Tensor, torch.Tensor]:
        v = self.cfg.vocab_size
        syn0 = torch.empty(v, self.K, dtype=torch.float32, device=self.dev)
        _C.init_syn0(syn0, self.rank * self.K, self.cfg.vector_size, int(seed))
        self.launches += 1
        syn1 = torch.zeros(v, self.K, dtype=torch.float32, device=self.dev)
        return syn0, syn1

    def upload_noise(self, alias, keep_thresh: np.ndarray):
        self._scale_window = -1               # the damping tables depend on the counts
        self._scale_cache = {}
        self.alias_dev = torch.from_numpy(alias.packed()).to(self.dev)
        self.keep_dev = torch.from_numpy(keep_thresh.view(np.int32).copy()).to(self.dev)
        self.subsample_active = bool((keep_thresh != U32_MAX).any())

    def release(self):
        self._xchg = None
        self._serve = None
        if self._nn is not None:
            self._nn.release()
        self._nn = None
        self.alias_dev = self.keep_dev = None
        self._cap = 0

    def _ensure_capacity(self, t: int):
        if t <= self._cap:
            return
        cap = max(t, 1024)
        d = self.dev
        if self._cap > 0:
            torch.cuda.synchronize(d)             # growing: earlier steps may still be using the old buffers
        # staging ring: pinned host buffer + device buffer + events per slot, so the host can run ahead of the
        # device (trainer.train queues up to 64 steps) without overwriting a pinned buffer whose H2D is pending
        self._stage = [_StageSlot(cap, d) for _ in range(N_STAGE)]
        self._stage_i = -1
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=d)
            self._pin_stats = torch.zeros(STATS_RING, 4, dtype=torch.float32).pin_memory()
            # events are created once: cudaEventCreate through torch costs ~0.4 ms each (measured in the fit loop)
            self._stats_ev = [torch.cuda.Event() for _ in range(STATS_RING)]
            self._stats_j = -1
        self.tok_c = torch.empty(cap, dtype=torch.int32, device=d)
        self.sid_c = torch.empty(cap, dtype=torch.int32, device=d)
        self.count = torch.zeros(1, dtype=torch.int32, device=d)
        self.sc_tiles = torch.zeros(int(_C.subsample_max_blocks(cap)) + 1, dtype=torch.int32, device=d)
        self._stats_ring = torch.zeros(256, 4, dtype=torch.float32, device=d)
        self._count_val = -1
        # pair-generation workspaces (pairgen.cu)
        cfg = self.cfg
        self.pd = int(_C.pairgen_desc_ints(cfg.negatives))
        self.pg_cinfo = torch.empty(cap, dtype=torch.int32, device=d)
        self.pg_off = torch.empty(cap, dtype=torch.int32, device=d)
        self.pg_npairs = torch.zeros(1, dtype=torch.int32, device=d)
        self.pg_tiles = torch.zeros(int(_C.pairgen_max_blocks(cap)) + 1, dtype=torch.int32, device=d)
        if self._tile_mode:
            # tensor-core tile kernel: no pair descriptors, only the shared negatives of every tile
            self.pg_desc = None
            self.tile_negs = torch.zeros(int(_C.sgns_tile_max_tiles(cap)) * cfg.tile_negatives, dtype=torch.int32, device=d)
        else:
            # n > 7 negatives: several descriptors per pair (csrc/pairgen.cu)
            self.pg_desc = torch.empty(cap * 2 * cfg.window * self.pd * int(_C.pairgen_splits(cfg.negatives)),
                                       dtype=torch.int32, device=d)
        self._cap = cap

    # ------------------------------------------------------------------ cross-shard exchange
    def _setup_exchange(self):
        """Symmetric exchange slots + flags for the in-kernel all-reduce (collective)."""
        from ..parallel.symm import alloc_symmetric
        cfg = self.cfg
        dev_index = self.dev.index or 0
        grid = int(_C.sgns_pairs_grid(self.K, dev_index, True))
        warps, nslot, slot_floats = [int(x) for x in _C.sgns_pairs_multi_geometry()]
        units = grid * warps                          # one exchange ring per warp
        xbytes = units * nslot * self.world * slot_floats * 4
        nseq = units
        fbytes = (units * self.world * 4 + 255) // 256 * 256
        xbytes = (xbytes + 255) // 256 * 256
        if self._loopback > 1:
            from ..parallel.symm import SymmBuffer
            local = torch.zeros(xbytes + fbytes, dtype=torch.uint8, device=self.dev)
            buf = SymmBuffer(local, [local.data_ptr()] * self.world, 0, None)
        else:
            # all ranks must launch the identical geometry (the flag protocol pairs warp w with warp w)
            g = torch.tensor([grid, -grid], dtype=torch.int64, device=self.dev)
            dist.all_reduce(g, op=dist.ReduceOp.MIN, group=self.e.comm.group)
            if int(g[0].item()) != grid or int(-g[1].item()) != grid:
                raise RuntimeError("ranks disagree on the persistent grid size; heterogeneous GPUs are not supported")
            buf = alloc_symmetric(xbytes + fbytes, self.dev, self.e.comm.group)
        self._pick_exchange_format(buf)
        self._xchg = {
            "buf": buf, "grid": grid, "slot_floats": slot_floats, "variant": "pairs",
            "xptrs": list(buf.ptrs), "fptrs": [p + xbytes for p in buf.ptrs],
            # NVLS multicast push (multimem.st) is opt-in: GW2V_NVLS=1 and a multicast mapping granted by the driver
            "mc_x": buf.multicast_ptr if (buf.multicast_ptr and self._want_nvls() and not (self.debug & 32)) else 0,
            "cta_seq": torch.zeros(nseq, dtype=torch.int32, device=self.dev),
            "err": torch.zeros(1, dtype=torch.int32, device=self.dev),
        }
        self.timing = torch.zeros(2, dtype=torch.int64, device=self.dev)

    def _pick_exchange_format(self, buf):
        """Exchange chunk format of the pair kernel (csrc/sgns_pairs.cu, PK_XSTRIDE): ``GW2V_XCHG=fast`` = 16-byte
        {f, f, f, tag} chunks, ``safe`` = 64-bit (value, tag) words, ``auto`` (default) = fast after a start-up self-test
        that streams 16-byte chunks between every pair of ranks and looks for torn reads (collective)."""
        mode = os.environ.get("GW2V_XCHG", "auto")
        if mode not in ("auto", "fast", "safe"):
            raise ValueError(f"GW2V_XCHG must be auto, fast or safe, not {mode!r}")
        self.xchg_selftest = None
        if mode == "auto" and self._loopback <= 1:
            res = torch.zeros(2, dtype=torch.int64, device=self.dev)
            buf.local[:self.world * 32 * 16].zero_()
            torch.cuda.synchronize(self.dev)
            dist.barrier(group=self.e.comm.group)
            _C.xchg_selftest(list(buf.ptrs), self.rank, int(os.environ.get("GW2V_XCHG_SELFTEST_ITERS", "200000")), res)
            torch.cuda.synchronize(self.dev)
            dist.all_reduce(res, op=dist.ReduceOp.SUM, group=self.e.comm.group)
            torn, seen = int(res[0].item()), int(res[1].item())
            buf.local[:self.world * 32 * 16].zero_()
            torch.cuda.synchronize(self.dev)
            dist.barrier(group=self.e.comm.group)
            self.xchg_selftest = {"torn": torn, "observed": seen}
            if torn:
                import logging
                logging.getLogger(__name__).warning(
                    "exchange self-test saw %d torn 16-byte chunks in %d reads: using 64-bit tagged words", torn, seen)
                mode = "safe"
        if mode == "safe":
            self.debug |= 32

    def _setup_tile_exchange(self):
        """Symmetric exchange ring of the tensor-core tile kernel (collective): per CTA ``slots`` x ``world`` payloads of
        128 centres x (band slots + tile negatives) partial dots, plus one release flag per (CTA, slot, source rank)."""
        from ..parallel.symm import SymmBuffer, alloc_symmetric
        cfg = self.cfg
        grid = self._tile_grid
        slots, fl = [int(v) for v in _C.sgns_tile_exchange_geometry(cfg.window, WINDOW_MODES[cfg.window_mode],
                                                                    cfg.tile_negatives)]
        xbytes = (grid * slots * self.world * fl * 4 + 255) // 256 * 256
        fbytes = (grid * slots * self.world * 4 + 255) // 256 * 256
        if self._loopback > 1:
            local = torch.zeros(xbytes + fbytes, dtype=torch.uint8, device=self.dev)
            buf = SymmBuffer(local, [local.data_ptr()] * self.world, 0, None)
        else:
            # every rank must run the identical persistent grid: CTA c of rank a exchanges with CTA c of rank b
            g = torch.tensor([grid, -grid], dtype=torch.int64, device=self.dev)
            dist.all_reduce(g, op=dist.ReduceOp.MIN, group=self.e.comm.group)
            if int(g[0].item()) != grid or int(-g[1].item()) != grid:
                raise RuntimeError("ranks disagree on the persistent grid size; heterogeneous GPUs are not supported")
            buf = alloc_symmetric(xbytes + fbytes, self.dev, self.e.comm.group)
        self._xchg = {
            "buf": buf, "grid": grid, "variant": "tile",
            "xptrs": list(buf.ptrs), "fptrs": [p + xbytes for p in buf.ptrs],
            "cta_seq": torch.zeros(grid, dtype=torch.int32, device=self.dev),
            "err": torch.zeros(1, dtype=torch.int32, device=self.dev),
        }
        self.timing = torch.zeros(2, dtype=torch.int64, device=self.dev)

    def _want_nvls(self) -> bool:
        """multimem.st push: transport="nvls" (or GW2V_NVLS=1) and a multicast mapping granted by the driver."""
        env = os.environ.get("GW2V_NVLS")
        return env == "1" if env is not None else self.e.opts.transport == "nvls"

    # ------------------------------------------------------------------ training
    def stage_tokens(self, tokens, sent_id):
        """Host -> device copy of one step's inputs on the copy stream (overlaps the previous step's kernels).
        Returns ``(slot, t)``.  The host blocks only when it is ``N_STAGE`` steps ahead of the device."""
        t = int(len(tokens))
        self._ensure_capacity(t)
        self._stage_i = (self._stage_i + 1) % N_STAGE
        slot = self._stage[self._stage_i]
        if slot.free_ev is not None:
            slot.free_ev.synchronize()            # previous user of this slot (H2D + kernels) has finished
        if isinstance(tokens, torch.Tensor) and tokens.is_pinned() and isinstance(sent_id, torch.Tensor) \
                and sent_id.is_pinned():
            src_tok, src_sid = tokens, sent_id
        else:
            slot.pin_tok_np[:t] = tokens.numpy() if isinstance(tokens, torch.Tensor) else tokens
            slot.pin_sid_np[:t] = sent_id.numpy() if isinstance(sent_id, torch.Tensor) else sent_id
            src_tok, src_sid = slot.pin_tok[:t], slot.pin_sid[:t]
        cs = self._copy_stream
        with torch.cuda.stream(cs):
            slot.tok[:t].copy_(src_tok, non_blocking=True)
            slot.sid[:t].copy_(src_sid, non_blocking=True)
            slot.h2d_ev.record(cs)
        return self._stage_i, t

    def train_step(self, tokens, sent_id, raw_pos0: int, iteration: int, alpha: float) -> torch.Tensor:
        si, t = self.stage_tokens(tokens, sent_id)
        return self.train_step_staged(si, t, raw_pos0, iteration, alpha)

    def train_step_staged(self, si: int, t: int, raw_pos0: int, iteration: int, alpha: float) -> torch.Tensor:
        slot = self._stage[si]
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(slot.h2d_ev)
        stats = self.train_step_device(slot.tok, slot.sid, t, raw_pos0, iteration, alpha)
        if slot.free_ev is None:
            slot.free_ev = slot._free_ev
        slot.free_ev.record(cur)
        return stats

    def train_step_async(self, tokens, sent_id, raw_pos0: int, iteration: int, alpha: float) -> "StepHandle":
        """Like ``train_step`` but the statistics come back through an asynchronous D2H copy into a pinned
        ring; ``handle.result()`` waits for THIS step only, so a caller that reads step s-1 after queueing
        step s keeps the device busy (H2D of s+1 and the read of s-1 both overlap the kernels of s)."""
        stats = self.train_step(tokens, sent_id, raw_pos0, iteration, alpha)
        self._stats_j = (self._stats_j + 1) % STATS_RING
        j = self._stats_j
        self._pin_stats[j].copy_(stats, non_blocking=True)
        self._stats_ev[j].record(torch.cuda.current_stream(self.dev))
        return StepHandle(self._pin_stats, self._stats_ev[j], j)

    def train_step_device(self, tok_dev: torch.Tensor, sid_dev: torch.Tensor, t: int, raw_pos0: int,
                          iteration: int, alpha: float) -> torch.Tensor:
        """One step on tokens that already live on the device (int32 tensors of length >= t).

        With ``GW2V_MAX_INFLIGHT=m > 0`` at most m steps are queued on the device (the host waits for the end of
        step s - m before queueing step s); the host needs ~35 us to queue a step, so the device never starves."""
        m = self._max_inflight
        if m > 0:
            ring = self._step_events
            i = self._step_i % m
            if ring[i] is not None:
                ring[i].synchronize()
            else:
                ring[i] = torch.cuda.Event()
        stats = self._train_step_device_impl(tok_dev, sid_dev, t, raw_pos0, iteration, alpha)
        if m > 0:
            self._step_events[self._step_i % m].record(torch.cuda.current_stream(self.dev))
            self._step_i += 1
        return stats

    def prepare(self, t: int):
        """One-time work of the first step of ``t`` tokens, callable ahead of the training loop (collective when
        world > 1): staging buffers, hot-row tables (numpy over the whole vocabulary) and the exchange rings."""
        self._ensure_capacity(t)
        self._update_row_scales(t)
        self._prepared_t = int(t)
        if self.world > 1 and self._xchg is None:
            if self._tile_mode:
                if not hasattr(self, "_tile_grid"):
                    self._tile_neg_scale = float(os.environ.get("GW2V_TILE_NEG_SCALE", self.e.tile_neg_scale()))
                    self._tile_neg_weight = float(os.environ.get("GW2V_TILE_NEG_WEIGHT", self.e.tile_neg_weight()))
                    self._tile_grid = int(os.environ.get("GW2V_TILE_GRID", self._props.multi_processor_count))
                self._setup_tile_exchange()
            else:
                self._setup_exchange()

    def warmup(self):
        """Launch one step over a single token with alpha = 0: every kernel of the step runs (and is loaded by the driver)
        without touching the weights (one token has no context, a zero-pair step is a no-op); collective like any step."""
        z = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._train_step_device_impl(z, z, 1, 0, 0, 0.0, scale_t=getattr(self, "_prepared_t", 1))

    def _update_row_scales(self, t: int):
        """Hot-row damping tables for a step of ``t`` tokens (models/engine.py::row_scales), cached per window size."""
        w = self.e.inflight_tokens(t)
        if w == self._scale_window:
            return
        # a short tail step has a smaller (power-of-two) window; the tables are cached per window so that alternating
        # between full and tail steps does not redo the numpy pass over the vocabulary (0.3 s at V = 10 M)
        self._scale_window = w
        if w not in self._scale_cache:
            sc = self.e.row_scales(w)
            if sc is None or sc[0].shape[0] == 0:
                self._scale_cache[w] = (None, None)
            else:
                self._scale_cache[w] = (torch.from_numpy(sc[0]).to(self.dev), torch.from_numpy(sc[1]).to(self.dev))
        self.row_scale0, self.row_scale1 = self._scale_cache[w]

    def _train_step_device_impl(self, tok_dev, sid_dev, t, raw_pos0, iteration, alpha, scale_t=None) -> torch.Tensor:
        cfg = self.cfg
        e = self.e
        self._ensure_capacity(t)
        self._update_row_scales(t if scale_t is None else scale_t)
        if self.world > 1 and self._xchg is None and not self._tile_mode:
            self._setup_exchange()
        if self.subsample_active:
            _C.subsample_compact(tok_dev, sid_dev, t, self.keep_dev, int(cfg.seed), int(iteration),
                                 int(raw_pos0), self.tok_c, self.sid_c, self.count, self.sc_tiles)
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
        wm = WINDOW_MODES[cfg.window_mode]
        if self._tile_mode:
            if not hasattr(self, "_tile_grid"):
                self._tile_neg_scale = float(os.environ.get("GW2V_TILE_NEG_SCALE", e.tile_neg_scale()))
                self._tile_neg_weight = float(os.environ.get("GW2V_TILE_NEG_WEIGHT", e.tile_neg_weight()))
                self._tile_grid = int(os.environ.get("GW2V_TILE_GRID", self._props.multi_processor_count))
            if self.world > 1 and self._xchg is None:
                self._setup_tile_exchange()
            x = self._xchg if self.world > 1 else None
            _C.sgns_step_tile(e.syn0, e.syn1, tok, sid, self.count, t, self.alias_dev, stats, int(raw_pos0),
                              int(cfg.seed), int(iteration), cfg.window, cfg.negatives, wm, float(alpha),
                              float(cfg.max_grad), self.compute_loss, self._tile_grid, self.debug, self.pg_cinfo,
                              self.pg_off, self.pg_npairs, self.pg_tiles, self.tile_negs, cfg.tile_negatives,
                              self.exp_table, self.row_scale0, self.row_scale1, self.tile_dbg,
                              self.world, self.rank, x["xptrs"] if x else [], x["fptrs"] if x else [],
                              x["cta_seq"] if x else None, x["err"] if x else None, self.timing if x else None,
                              self._tile_neg_scale, self._tile_neg_weight)
            self.launches += 4            # pair_count, pair_tile_scan, tile_negs, sgns_tile
            return stats
        # pair_count, pair_tile_scan (zeroes the statistics), pair_fill, then the training kernel
        if self.world > 1:
            x = self._xchg
            _C.sgns_step_pairs(e.syn0, e.syn1, tok, sid, self.count, t, self.alias_dev, stats, int(raw_pos0),
                               int(cfg.seed), int(iteration), cfg.window, cfg.negatives, wm, float(alpha),
                               float(cfg.max_grad), self.compute_loss, x["grid"], self.world, self.rank,
                               x["xptrs"], x["fptrs"], x["cta_seq"], x["err"], self.timing, self.debug,
                               self.pg_cinfo, self.pg_off, self.pg_npairs, self.pg_desc, self.pg_tiles,
                               x["mc_x"], self._share_centre, self.exp_table, self.row_scale0, self.row_scale1)
        else:
            if not hasattr(self, "_grid1"):
                self._grid1 = int(_C.sgns_pairs_grid(self.K, self.dev.index or 0, False))
            _C.sgns_step_pairs(e.syn0, e.syn1, tok, sid, self.count, t, self.alias_dev, stats, int(raw_pos0),
                               int(cfg.seed), int(iteration), cfg.window, cfg.negatives, wm, float(alpha),
                               float(cfg.max_grad), self.compute_loss, self._grid1, 1, 0, [], [], None, None,
                               None, self.debug, self.pg_cinfo, self.pg_off, self.pg_npairs, self.pg_desc,
                               self.pg_tiles, 0, self._share_centre, self.exp_table, self.row_scale0,
                               self.row_scale1)
        self.launches += 4
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

    # ------------------------------------------------------------------ inference
    def serve(self):
        """Lazily allocated symmetric serving workspace (collective on first use)."""
        if self._serve is None:
            from .serving import ServeExchange
            self._serve = ServeExchange(self)
        return self._serve

    def nn_index(self):
        """Lazily built nearest-neighbour index (row-sharded replica + fused select kernel, ops/nn.py)."""
        if self._nn is None:
            from .nn import NNIndex
            self._nn = NNIndex(self)
        return self._nn

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
