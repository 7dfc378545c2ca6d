"""Per-rank shard engine: the B200 counterpart of a Glint parameter server.

One ``ShardEngine`` lives in each process (one process per GPU).  It owns the
rank's column slice of ``syn0`` (input vectors, "u") and ``syn1neg`` (output
vectors, "v") for the FULL vocabulary plus the replicated noise tables, and
implements every server-side operation the reference calls on
``BigWord2VecMatrix``:

====================  =========================================================
reference (Glint) op  here
====================  =========================================================
``dotprod``+``adjust``  ``train_step`` (one fused sm_100a kernel incl. the
                        cross-shard partial-dot all-reduce; CPU: torch ops +
                        ``all_reduce``) -- MLLIB:421-425
``pull``               ``pull``            -- MLLIB:514,539,639,652
``pullAverage``        ``pull_average``    -- ML:453
``norms``              ``norms``           -- MLLIB:486
``multiply``           ``multiply`` / ``top_k`` (scores + cosine + top-k fused
                        on device)         -- MLLIB:598-617
``save`` / load        ``save_shard`` / ``load_columns``  -- MLLIB:494,722
``destroy``            ``destroy``         -- MLLIB:665
====================  =========================================================
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from ..data.sampler import AliasTable, keep_thresholds, unigram_alias, unigram_table_alias
from ..parallel.comm import Comm
from ..parallel.sharding import ColumnShard, make_shard
from . import sgns
from .sgns import SGNSConfig


@dataclass
class EngineOptions:
    """Engine knobs carried by ``parameterServerConfig``."""
    batch_size: int = 50            # centres per mini-batch (reference ``batchSize``)
    step_tokens: int = 0            # tokens per device step (0 = auto)
    subsample_ratio: float = 1e-6
    subsample_mode: str = "reference"   # "reference" (Q1: the reference's sub-sampling is inert) | "word2vec"
    max_hot_updates: int = 256          # (damping off) auto step size: expected stale summed updates on the hottest row
    # Hot-row damping: a device step applies the summed, stale updates of thousands of centres at once; the rows of very
    # frequent words then receive hundreds of them and blow up (README.md:17-19 warns about exactly this; a 50-centre
    # mini-batch of the reference is 10^3 times smaller).  The update of row r is scaled by min(1, hot_row_cap / c_r),
    # c_r = expected number of concurrent updates of the row inside one staleness window -- an adaptive per-row
    # learning rate that leaves all but the most frequent rows untouched.  Applied by the fused GPU kernels (pairs and
    # tile); the un-fused path keeps the reference's exact semantics.  0 = off.
    hot_row_cap: float = 32.0
    # neg_sharing="tile": weight of the negative term in the UPDATES, both the dU and the dV side (1 = the reference's n
    # negatives; 0.4 behaves like n = 2).  A stability knob for constant-learning-rate runs with very large norms; the
    # loss statistic always uses the full weight.  Scaling only one side biases the model (profiles/r2_tile_kernel.md).
    tile_neg_weight: float = 1.0
    # asynchronous data-parallel workers of the reference (``numPartitions``, MLLIB:122-126,345,392): P workers each
    # have one mini-batch in flight against the same servers, i.e. the updates of P * batchSize centres are computed
    # from the same stale rows.  The un-fused engine path reproduces that staleness: mini-batches of
    # ``batch_size * num_partitions`` centres ("Use a small number for accuracy", MLLIB:120).  The GPU kernels always
    # have thousands of centres in flight; there the value only enters the damping estimate.
    num_partitions: int = 1
    # negative sampler: "alias" = exact cn^0.75 through a Vose alias table (8 bytes per word); "table" = the quantised
    # distribution of the reference's ``unigramTableSize``-slot table (MLLIB:239-244, ML:204-209), sampled through the
    # same alias machinery (data/sampler.py::unigram_table_counts) -- a parity mode, no 400 MB table is built
    sampler: str = "alias"
    unigram_table_size: int = 100_000_000
    # How partial dots travel between column shards on GPUs:
    #   "auto"/"p2p": in-kernel st.global pushes into the peers' symmetric memory (the product)
    #   "nvls":       same kernel, one multimem.st per chunk on the NVLS multicast mapping
    #   "nccl"/"gloo": the un-fused Glint-style path (dotprod -> library all-reduce -> adjust) with the reference's
    #                  mini-batch semantics on any device; slow, kept for A/B runs and as the CPU path
    transport: str = "auto"
    # training kernel: the pair kernel (csrc/sgns_pairs.cu) for neg_sharing pair/centre, the tcgen05 tile kernel
    # (csrc/sgns_tile.cu) for neg_sharing="tile"; "auto" = that rule.  The round-1 generations live in csrc/legacy.
    kernel: str = "auto"
    store_syn1: bool = True         # keep syn1neg in saves (retrainable)

    @classmethod
    def from_dict(cls, d: Optional[dict]):
        d = dict(d or {})
        known = {k: d[k] for k in list(d) if k in cls.__dataclass_fields__}
        return cls(**known)

    def __post_init__(self):
        if self.transport not in ("auto", "p2p", "nvls", "nccl", "gloo"):
            raise ValueError(f"unknown transport {self.transport!r}")
        if self.kernel not in ("auto", "pairs", "tile"):
            raise ValueError(f"unknown kernel {self.kernel!r} (the group/pipe/v1 kernels of round 1 are retired)")
        if self.sampler not in ("alias", "table"):
            raise ValueError(f"unknown sampler {self.sampler!r}")


def round_window(w: int) -> int:
    """Staleness window rounded UP to 4 significant bits: at most 12.5 % more damping than the exact window, and only a
    handful of distinct values per octave for the table cache."""
    w = max(1, int(w))
    q = 1 << max(0, w.bit_length() - 4)
    return (w + q - 1) // q * q


def _host_i32(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        x = x.cpu().numpy()
    return np.asarray(x, dtype=np.int32)


class _ReadyHandle:
    def __init__(self, value):
        self._value = value

    def result(self):
        return torch.as_tensor(self._value)


class ShardEngine:
    def __init__(self, cfg: SGNSConfig, comm: Optional[Comm] = None,
                 device: Optional[torch.device] = None, options: Optional[EngineOptions] = None):
        self.cfg = cfg
        self.comm = comm if comm is not None else Comm()
        self.opts = options if options is not None else EngineOptions()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
                else torch.device("cpu")
        self.device = torch.device(device)
        self.shard: ColumnShard = make_shard(cfg.vector_size, self.comm.world, self.comm.rank)
        self.syn0: Optional[torch.Tensor] = None      # [V, K] fp32
        self.syn1: Optional[torch.Tensor] = None
        self.alias: Optional[AliasTable] = None
        self.keep_thresh: Optional[np.ndarray] = None
        self.noise_counts: Optional[np.ndarray] = None
        self._norms: Optional[torch.Tensor] = None
        self._version = 0                    # bumped whenever syn0 may have changed (serving caches key on it)
        self._cuda = None
        if self.device.type == "cuda":
            from ..ops import cuda as _cuda_ops      # raises loudly if the extension is missing
            self._cuda = _cuda_ops.CudaShardOps(self)

    # ------------------------------------------------------------------ setup
    def _touch(self):
        """syn0 changed (or may have): drop the cached norms and every serving replica built from it."""
        self._norms = None
        self._version += 1

    @property
    def is_cuda(self) -> bool:
        return self.device.type == "cuda"

    @property
    def unfused(self) -> bool:
        """True when the step runs as dotprod -> library all-reduce -> adjust: on CPUs and with transport nccl/gloo
        (neg_sharing="tile" then uses library GEMMs; on GPUs it is the tcgen05 kernel csrc/sgns_tile.cu)."""
        return (not self.is_cuda) or self.opts.transport in ("nccl", "gloo")

    @property
    def vocab_size(self) -> int:
        return self.cfg.vocab_size

    def init_weights(self, seed: Optional[int] = None):
        """``syn0 ~ U(-0.5,0.5)/d`` on this shard's columns, ``syn1neg = 0``."""
        seed = self.cfg.seed if seed is None else seed
        sh = self.shard
        v = self.cfg.vocab_size
        if self.is_cuda:
            self.syn0, self.syn1 = self._cuda.init_weights(seed)
        else:
            # scale uses the logical d; columns >= d are zero padding
            full0, _ = sgns.init_embeddings(v, sh.padded_vector_size, seed, scale_dim=self.cfg.vector_size)
            full0[:, self.cfg.vector_size:] = 0
            self.syn0 = full0[:, sh.rank * sh.cols:(sh.rank + 1) * sh.cols].contiguous()
            self.syn1 = torch.zeros(v, sh.cols, dtype=torch.float32)
        self._touch()

    def set_noise(self, counts: np.ndarray, use_native: bool = True):
        """Build the unigram^0.75 alias table and the sub-sampling thresholds
        from the vocabulary counts (``bcVocabCns``, MLLIB:317,355)."""
        counts = np.asarray(counts, dtype=np.int64)
        if counts.shape[0] != self.cfg.vocab_size:
            raise ValueError("counts length != vocab size")
        self.noise_counts = counts
        if self.opts.sampler == "table":
            self.alias = unigram_table_alias(counts, int(self.opts.unigram_table_size), 0.75, use_native=use_native)
        else:
            self.alias = unigram_alias(counts, 0.75, use_native=use_native)
        self.keep_thresh = keep_thresholds(counts, self.opts.subsample_ratio, self.opts.subsample_mode)
        if self.is_cuda:
            self._cuda.upload_noise(self.alias, self.keep_thresh)

    def inflight_tokens(self, step_tokens: int) -> int:
        """Staleness window: how many centres have their updates computed from the same (stale) rows."""
        if not self.is_cuda or self.unfused:
            return max(1, self.opts.batch_size) * max(1, self.opts.num_partitions)
        if self.cfg.neg_sharing == "tile":
            w = min(int(step_tokens), 148 * self.cfg.tile_centres)         # one tile per SM in flight
        else:
            w = min(int(step_tokens), 4096)                                 # resident warps x pairs of the pair kernel
        # rounded up to 4 significant bits (<= 12.5 % over): short tail steps then share the damping tables of a few
        # window sizes instead of each paying a numpy pass over the vocabulary (ops/cuda.py::_update_row_scales caches
        # per window).  Rounding to a power of two was tried first and cost quality: up to 1.75x more damping
        # (smoke corpus: planted recall 0.95 -> 0.65).
        return round_window(w)

    def mean_pairs_per_centre(self) -> float:
        w = self.cfg.window
        if self.cfg.window_mode == "reference":
            return float(np.mean([max(0, 2 * b - 1) for b in range(w)]))
        return float(w + 1)

    def tile_neg_scale(self) -> float:
        """neg_sharing="tile": scale of the shared negatives' row updates.  One tile adds the summed, stale contributions
        of all its centres to each shared negative row in one event of mass ``tile_centres * m * n / tile_negatives``
        unit updates (~90 for 128 centres, 32 negatives, window 5) -- beyond ``hot_row_cap`` it is capped like a hot row
        (measured: without this the 10 M x 512 benchmark run diverges in tile mode, max|dot| ~ 8e5)."""
        cap = float(self.opts.hot_row_cap)
        if cap <= 0 or self.cfg.neg_sharing != "tile":
            return 1.0
        mass = self.cfg.tile_centres * self.mean_pairs_per_centre() * self.cfg.negatives / self.cfg.tile_negatives
        return float(min(1.0, cap / max(mass, 1e-30)))

    def tile_neg_weight(self) -> float:
        """neg_sharing="tile": weight of the negative term in the updates (both sides; 1 = the reference's n)."""
        return float(self.opts.tile_neg_weight)

    def row_scales(self, window_tokens: int):
        """Hot-row damping tables ``(scale0, scale1)`` (float32, length = number of hot rows H; rows >= H are 1).

        Expected concurrent updates of row r inside a window of W tokens (after sub-sampling):
        ``c0 = W f_r m`` for the centre row, ``c1 = W m (f_r + n q_r)`` for the output row (context + negative
        draws), with f the effective token frequency, q the noise distribution (cn^0.75) and m the mean number of
        pairs per centre of the window mode; scale = min(1, hot_row_cap / c)."""
        cap = float(self.opts.hot_row_cap)
        if cap <= 0 or self.noise_counts is None:
            return None
        cnt = self.noise_counts.astype(np.float64)
        keep = (self.keep_thresh.astype(np.float64) + 1.0) / 2.0 ** 32
        eff = cnt * np.minimum(keep, 1.0)
        f = eff / max(eff.sum(), 1.0)
        q = cnt ** 0.75
        q /= max(q.sum(), 1e-300)
        m = self.mean_pairs_per_centre()
        W = float(max(1, window_tokens))
        c0 = W * f * m
        c1 = W * m * (f + self.cfg.negatives * q)
        s0 = np.minimum(1.0, cap / np.maximum(c0, 1e-30)).astype(np.float32)
        s1 = np.minimum(1.0, cap / np.maximum(c1, 1e-30)).astype(np.float32)
        hot = np.nonzero((s0 < 1.0) | (s1 < 1.0))[0]
        h = int(hot.max()) + 1 if hot.size else 0
        return s0[:h], s1[:h]

    def set_weights(self, syn0_full: Optional[torch.Tensor], syn1_full: Optional[torch.Tensor] = None):
        """Install this rank's column slice of full [V, d] matrices."""
        sh = self.shard
        v = self.cfg.vocab_size

        def _slice(full):
            out = torch.zeros(v, sh.cols, dtype=torch.float32)
            if sh.real_cols > 0:
                out[:, :sh.real_cols] = full[:, sh.col_start:sh.col_start + sh.real_cols].to(torch.float32)
            return out.to(self.device)

        if syn0_full is not None:
            self.syn0 = _slice(syn0_full)
        self.syn1 = _slice(syn1_full) if syn1_full is not None else \
            torch.zeros(v, sh.cols, dtype=torch.float32, device=self.device)
        self._touch()

    def destroy(self):
        self.syn0 = self.syn1 = None
        self._touch()
        if self._cuda is not None:
            self._cuda.release()

    # --------------------------------------------------------------- training
    def train_step(self, tokens, sent_id, raw_pos0: int, iteration: int, alpha: float):
        """One device step over ``tokens`` (whole sentences).

        Sub-sampling, windowing, negative sampling, partial dots, the
        cross-shard all-reduce, sigmoid/LR and the row updates all happen here.
        Returns a 4-vector ``[pairs, loss, max|dot|, kept_tokens]`` (a device
        tensor on GPUs so the caller decides when to synchronise).
        """
        if self.alias is None:
            raise RuntimeError("set_noise() must be called before training")
        self._touch()
        if self.is_cuda and not self.unfused:
            return self._cuda.train_step(tokens, sent_id, raw_pos0, iteration, alpha)
        return self._train_step_cpu(_host_i32(tokens), _host_i32(sent_id), raw_pos0, iteration, alpha)

    def train_step_async(self, tokens, sent_id, raw_pos0: int, iteration: int, alpha: float):
        """``train_step`` whose statistics are read back asynchronously: returns a handle with
        ``result() -> CPU tensor``.  On GPUs the host -> device copy of the NEXT step and the read-back of
        the PREVIOUS one overlap this step's kernels (``ops/cuda.py::train_step_async``)."""
        if self.alias is None:
            raise RuntimeError("set_noise() must be called before training")
        self._touch()
        if self.is_cuda and not self.unfused:
            return self._cuda.train_step_async(tokens, sent_id, raw_pos0, iteration, alpha)
        return _ReadyHandle(self._train_step_cpu(_host_i32(tokens), _host_i32(sent_id), raw_pos0, iteration, alpha))

    def _train_step_cpu(self, tokens, sent_id, raw_pos0, iteration, alpha):
        cfg = self.cfg
        keep = sgns.subsample_mask(tokens, self.keep_thresh, cfg.seed, iteration, raw_pos0)
        tokens = tokens[keep]
        sent_id = sent_id[keep]
        t = tokens.shape[0]
        pairs = 0
        loss = 0.0
        maxdot = 0.0
        # async workers = staleness window.  This path IS the reference's semantics (no damping: like the reference it
        # relies on small mini-batches, "Use a small number [of partitions] for accuracy", MLLIB:120)
        bs = max(1, self.opts.batch_size) * max(1, self.opts.num_partitions)
        for lo in range(0, t, bs):
            st = self._minibatch_cpu(tokens, sent_id, raw_pos0, iteration, alpha, lo, min(t, lo + bs))
            pairs += st.pairs
            loss += st.loss
            maxdot = max(maxdot, st.max_abs_dot)
        return torch.tensor([pairs, loss, maxdot, t], dtype=torch.float64)

    # --- the unfused Glint-style API (dotprod / adjust), used on CPU and by the baseline
    def partial_dots(self, w: torch.Tensor, c: torch.Tensor, ng: torch.Tensor) -> torch.Tensor:
        """Partial dot products over this shard's columns: [P, 1+n]."""
        u = self.syn0[w]
        f = torch.empty(w.shape[0], 1 + ng.shape[1], dtype=torch.float32, device=self.device)
        f[:, 0] = (u * self.syn1[c]).sum(-1)
        f[:, 1:] = torch.einsum("pd,pnd->pn", u, self.syn1[ng])
        return f

    def adjust(self, w, c, ng, gplus, gminus):
        """Row updates on this shard's columns from pre-update values."""
        u = self.syn0[w]
        vc = self.syn1[c]
        vn = self.syn1[ng]
        du = gplus[:, None] * vc + torch.einsum("pn,pnd->pd", gminus, vn)
        self.syn1.index_add_(0, c, gplus[:, None] * u)
        self.syn1.index_add_(0, ng.reshape(-1), (gminus[:, :, None] * u[:, None, :]).reshape(-1, u.shape[1]))
        self.syn0.index_add_(0, w, du)

    def _minibatch_cpu(self, tokens, sent_id, pos0, iteration, alpha, lo, hi) -> sgns.StepStats:
        cfg = self.cfg
        ci, cj, slot = sgns.enumerate_pairs(cfg, tokens, sent_id, pos0, iteration, lo, hi)
        stats = sgns.StepStats(pairs=int(ci.shape[0]))
        # every rank enumerates the same pairs, so an empty batch is empty everywhere
        if ci.shape[0] == 0:
            return stats
        if cfg.neg_sharing == "tile":
            return self._minibatch_tile(tokens, pos0, iteration, alpha, ci, cj, stats)
        pos = np.uint64(pos0) + ci.astype(np.uint64)
        neg = sgns.draw_negatives(cfg, self.alias, pos, slot, iteration)
        tok = tokens.astype(np.int64)
        w = torch.from_numpy(tok[ci]).to(self.device)
        c = torch.from_numpy(tok[cj]).to(self.device)
        ng = torch.from_numpy(neg.astype(np.int64)).to(self.device)
        f = self.partial_dots(w, c, ng)
        f = self.comm.all_reduce_sum(f)                   # the Glint client-side aggregation
        neg_mask = (ng != c[:, None]).to(torch.float32)
        gplus = sgns.sigmoid_coeff(f[:, 0], 1.0, alpha, cfg.sigmoid_mode, cfg.max_grad)
        gminus = sgns.sigmoid_coeff(f[:, 1:], 0.0, alpha, cfg.sigmoid_mode, cfg.max_grad) * neg_mask
        stats.loss = float(sgns.sgns_loss(f[:, 0], f[:, 1:], neg_mask))
        stats.max_abs_dot = float(f.abs().max())
        self.adjust(w, c, ng, gplus, gminus)
        return stats

    def _minibatch_tile(self, tokens, pos0, iteration, alpha, ci, cj, stats) -> sgns.StepStats:
        """neg_sharing="tile" on a column shard with library GEMMs (docs/round2_tile_gemm.md): per tile
        ``S_neg = U @ Vneg^T`` (partial over this shard's columns) -> one all-reduce of all partial dots of the
        mini-batch -> coefficients -> ``dU = G @ Vneg``, ``dVneg = G^T @ U`` and the per-pair positive updates."""
        cfg, dev = self.cfg, self.device
        tok = tokens.astype(np.int64)
        w = torch.from_numpy(tok[ci]).to(dev)
        c = torch.from_numpy(tok[cj]).to(dev)
        centres, m, tile = sgns.tile_terms(cfg, ci)
        tiles, first = np.unique(tile, return_index=True)
        bounds = np.append(first, len(centres))                        # active centres of tile k: [bounds[k], bounds[k+1])
        tneg = sgns.tile_negatives(cfg, self.alias, pos0, tiles, iteration)
        wa = torch.from_numpy(tok[centres]).to(dev)
        u, vc, ua = self.syn0[w], self.syn1[c], self.syn0[wa]
        nn = cfg.tile_negatives
        part = torch.empty(ci.shape[0] + len(centres) * nn, dtype=torch.float32, device=dev)
        part[:ci.shape[0]] = (u * vc).sum(-1)
        fneg = part[ci.shape[0]:].view(len(centres), nn)
        negs = [torch.from_numpy(tneg[k].astype(np.int64)).to(dev) for k in range(len(tiles))]
        for k in range(len(tiles)):                                    # GEMM 1 per tile
            a, b = int(bounds[k]), int(bounds[k + 1])
            fneg[a:b] = ua[a:b] @ self.syn1[negs[k]].t()
        full = self.comm.all_reduce_sum(part)                          # the Glint client-side aggregation
        fplus, fminus = full[:ci.shape[0]], full[ci.shape[0]:].view(len(centres), nn)
        wgt = torch.from_numpy(m.astype(np.float64) * cfg.negatives / nn).to(torch.float32).to(dev)
        gplus = sgns.sigmoid_coeff(fplus, 1.0, alpha, cfg.sigmoid_mode, cfg.max_grad)
        # tile_neg_weight scales the negative term of the UPDATES (both sides); the loss keeps the full weight
        gminus = sgns.sigmoid_coeff(fminus, 0.0, alpha, cfg.sigmoid_mode, cfg.max_grad) * (wgt * float(self.opts.tile_neg_weight))[:, None]
        stats.loss = float(sgns.sgns_loss(fplus, fminus, wgt[:, None].expand_as(fminus)))
        stats.max_abs_dot = float(full.abs().max())
        du_neg = torch.empty_like(ua)
        for k in range(len(tiles)):                                    # GEMMs 2 and 3 per tile, pre-update rows
            a, b = int(bounds[k]), int(bounds[k + 1])
            vn = self.syn1[negs[k]]
            du_neg[a:b] = gminus[a:b] @ vn
            negs[k] = (negs[k], gminus[a:b].t() @ ua[a:b])
        self.syn1.index_add_(0, c, gplus[:, None] * u)
        for idx, dvn in negs:
            self.syn1.index_add_(0, idx, dvn)
        self.syn0.index_add_(0, w, gplus[:, None] * vc)
        self.syn0.index_add_(0, wa, du_neg)
        return stats

    # -------------------------------------------------------------- inference
    def _check_rows(self, rows: torch.Tensor) -> torch.Tensor:
        """Row indices are validated on the host: the device kernels do not bounds-check (a bad index on a GPU
        is an illegal address that kills the context, on the CPU an IndexError)."""
        if rows.numel() and (int(rows.min()) < 0 or int(rows.max()) >= self.cfg.vocab_size):
            raise IndexError(f"row index out of range [0, {self.cfg.vocab_size})")
        return rows

    def pull(self, rows) -> torch.Tensor:
        """Full [R, d] input vectors for ``rows`` (collective)."""
        rows = self._check_rows(torch.as_tensor(rows, dtype=torch.int64))
        if self.is_cuda and self._cuda.serve_fused:
            # the column all-gather happens inside the gather kernel (peer stores over NVLink)
            return self._cuda.serve().pull(self._cuda._rows_dev(rows))[:, :self.cfg.vector_size]
        if self.is_cuda:
            part = self._cuda.gather_rows(rows)
        else:
            part = self.syn0[rows]
        full = self.comm.all_gather_cols(part)
        return full[:, :self.cfg.vector_size]

    def pull_average(self, rows_flat, offsets) -> torch.Tensor:
        """Per-sentence mean of input vectors (empty sentence -> zeros), [S, d]."""
        rows_flat = self._check_rows(torch.as_tensor(rows_flat, dtype=torch.int64))
        offsets = torch.as_tensor(offsets, dtype=torch.int64)
        if offsets.numel() < 1 or int(offsets[0]) != 0 or int(offsets[-1]) != rows_flat.numel() \
                or bool((offsets[1:] < offsets[:-1]).any()):
            raise ValueError("offsets must be non-decreasing, start at 0 and end at len(rows_flat)")
        ns = offsets.shape[0] - 1
        if self.is_cuda and self._cuda.serve_fused:
            full = self._cuda.serve().pull_average(self._cuda._rows_dev(rows_flat), self._cuda._rows_dev(offsets))
            return full[:, :self.cfg.vector_size]
        if self.is_cuda:
            part = self._cuda.segment_mean_rows(rows_flat, offsets)
        else:
            part = torch.zeros(ns, self.shard.cols, dtype=torch.float32)
            lens = (offsets[1:] - offsets[:-1])
            if rows_flat.numel() > 0:
                seg = torch.repeat_interleave(torch.arange(ns), lens)
                part.index_add_(0, seg, self.syn0[rows_flat])
            part = part / lens.clamp(min=1).to(torch.float32)[:, None]
        full = self.comm.all_gather_cols(part)
        return full[:, :self.cfg.vector_size]

    def norms(self) -> torch.Tensor:
        """Euclidean norm of every input vector, [V] (cached; collective)."""
        if self._norms is None and self.is_cuda and self._cuda.serve_fused:
            # reduce-scatter + sqrt + all-gather inside the kernels (ops/serving.py)
            self._norms = self._cuda.serve().norms()
        if self._norms is None:
            if self.is_cuda:
                sq = self._cuda.row_sqnorm()
            else:
                sq = (self.syn0 * self.syn0).sum(-1)
            sq = self.comm.all_reduce_sum(sq)
            self._norms = sq.sqrt()
        return self._norms

    def multiply(self, q) -> torch.Tensor:
        """``syn0 @ q`` for a full-length query vector, [V] (collective)."""
        q = torch.as_tensor(q, dtype=torch.float32).reshape(1, -1)
        return self._scores(q)[0]

    def _query_slice(self, q: torch.Tensor) -> torch.Tensor:
        sh = self.shard
        out = torch.zeros(q.shape[0], sh.cols, dtype=torch.float32)
        if sh.real_cols:
            out[:, :sh.real_cols] = q[:, sh.col_start:sh.col_start + sh.real_cols]
        return out.to(self.device)

    def _scores(self, q: torch.Tensor) -> torch.Tensor:
        qs = self._query_slice(q)
        if self.is_cuda and self._cuda.serve_fused:
            sx = self._cuda.serve()
            return torch.stack([sx.multiply(qs[i:i + 1]) for i in range(qs.shape[0])], 0)
        if self.is_cuda:
            part = self._cuda.scores(qs)              # [Q, V]
        else:
            part = qs @ self.syn0.t()
        return self.comm.all_reduce_sum(part)

    def top_k(self, queries, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Cosine top-k against all input vectors (MLLIB:589-617).

        ``queries`` [Q, d] need not be normalised.  Zero-norm rows score 0.
        Returns (indices [Q, k], similarities [Q, k]) on the CPU.
        """
        q = torch.as_tensor(queries, dtype=torch.float32).reshape(-1, self.cfg.vector_size)
        if q.shape[0] == 0:                                   # no queries: nothing to launch (and no collective)
            kk = min(k, self.cfg.vocab_size)
            return torch.empty(0, kk, dtype=torch.int64), torch.empty(0, kk, dtype=torch.float32)
        qn = q.norm(dim=1, keepdim=True)
        q = torch.where(qn > 0, q / qn.clamp(min=1e-30), q)       # snrm2 / sscal (MLLIB:593-595)
        k = min(k, self.cfg.vocab_size)
        if self.is_cuda:
            # large vocabularies: row-sharded replica + score GEMM with in-epilogue selection (ops/nn.py); the dense
            # [Q, V] path below serves small vocabularies, huge k and the (never yet seen) candidate overflow
            nn = self._cuda.nn_index()
            if nn.supported(q.shape[0], k):
                res = nn.top_k(q, k)
                if res is not None:
                    return res[0].cpu(), res[1].cpu()
        norms = self.norms()
        if self.is_cuda:
            idx, sim = self._cuda.top_k(self._query_slice(q), norms, k)
            return idx.cpu(), sim.cpu()
        scores = self._scores(q)
        inv = torch.where(norms > 0, 1.0 / norms.clamp(min=1e-30), torch.zeros_like(norms))
        cos = scores * inv[None, :]
        sim, idx = torch.topk(cos, k, dim=1)
        return idx, sim

    # ------------------------------------------------------------ persistence
    def shard_tensors(self) -> Dict[str, torch.Tensor]:
        """This rank's real (un-padded) column slices as tensor views on the engine's device (no copy)."""
        sh = self.shard
        out = {"syn0": self.syn0[:, :sh.real_cols].detach()}
        if self.syn1 is not None and self.opts.store_syn1:
            out["syn1"] = self.syn1[:, :sh.real_cols].detach()
        return out

    def load_columns(self, name: str, col_start: int, block: np.ndarray):
        """Install saved columns ``[col_start, col_start + block.shape[1])`` of
        matrix ``name`` -- only the part overlapping this shard is kept, which
        is what makes loading with a different shard count work (Q12)."""
        sh = self.shard
        v = self.cfg.vocab_size
        for attr in ("syn0", "syn1"):
            if getattr(self, attr) is None:
                setattr(self, attr, torch.zeros(v, sh.cols, dtype=torch.float32, device=self.device))
        target = self.syn0 if name == "syn0" else self.syn1
        lo = max(col_start, sh.col_start)
        hi = min(col_start + block.shape[1], sh.col_start + sh.real_cols)
        if hi > lo:
            # row chunks: the (usually memory-mapped) file never has to fit in host memory at once
            step = max(1, (256 << 20) // ((hi - lo) * 4))
            for r0 in range(0, v, step):
                r1 = min(v, r0 + step)
                src = torch.from_numpy(np.array(block[r0:r1, lo - col_start:hi - col_start], dtype=np.float32, order="C"))
                target[r0:r1, lo - sh.col_start:hi - sh.col_start] = src.to(self.device)
        self._touch()
