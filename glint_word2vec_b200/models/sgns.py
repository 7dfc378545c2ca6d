"""Skip-gram with negative sampling: the step specification and its CPU oracle.

This file is the executable specification every device kernel is tested
against (SURVEY.md 4.3 "oracle" tier, Appendix B).  It realises the reference
training semantics (MLLIB:371-433 client side, Glint ``dotprod``/``adjust``
server side [G]):

* per (centre, context) pair ``n`` private negatives drawn from ``cn^0.75``;
* one mini-batch = all dot products from PRE-update weights, summed updates;
* ``g = (label - sigmoid(f)) * alpha`` with the hard clip at +-6 (MLLIB:292-302);
* reference window (radius ``b in [0, window-1]``, contexts ``i-b .. i+b-1``,
  MLLIB:385-386, quirk Q2) or word2vec.c window.

All randomness comes from ``utils.philox`` keyed by the token position, so the
oracle, the CPU engine and the CUDA kernels draw identical windows/negatives.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Optional, Tuple

import numpy as np
import torch

from ..data.sampler import AliasTable
from ..utils import philox

MAX_EXP = 6.0
EXP_TABLE_SIZE = 1000


@dataclass
class SGNSConfig:
    vocab_size: int
    vector_size: int = 100
    window: int = 5
    negatives: int = 5
    seed: int = 1
    window_mode: str = "reference"      # "reference" (Q2) | "word2vec_c"
    sigmoid_mode: str = "exact"         # "exact" | "table" (MLLIB:281-302 parity)
    max_grad: float = 0.0               # optional |g| clip (0 = off)
    # "pair": n private negatives per (centre, context) pair - the reference's behaviour (one seed per request,
    # MLLIB:420-421).  "centre": the n negatives are drawn once per centre and shared by all of its pairs
    # (pWord2Vec-style sharing): same expected gradient, the negative rows of a centre stay hot in L2.
    # "tile": ``tile_negatives`` negatives are drawn once per tile of ``tile_centres`` consecutive centres and
    # shared by all of them; the negative term of centre i is weighted m_i * n / tile_negatives (m_i = number of
    # its contexts) so that every pair still sees n negatives in expectation.  This turns the step into GEMMs
    # (docs/round2_tile_gemm.md); implemented by the oracle and the un-fused library path, not yet by a kernel.
    neg_sharing: str = "pair"
    tile_centres: int = 128
    tile_negatives: int = 64         # 64: half the per-tile event mass of 32 and twice the distinct negatives per token (profiles/r2_tile_kernel.md)

    def __post_init__(self):
        if self.neg_sharing not in ("pair", "centre", "tile"):
            raise ValueError(f"unknown neg_sharing {self.neg_sharing!r}")
        if self.tile_centres < 1 or self.tile_negatives < 1:
            raise ValueError("tile_centres and tile_negatives must be positive")
        if self.window_mode not in ("reference", "word2vec_c"):
            raise ValueError(f"unknown window_mode {self.window_mode!r}")
        if self.sigmoid_mode not in ("exact", "table"):
            raise ValueError(f"unknown sigmoid_mode {self.sigmoid_mode!r}")

    @property
    def ctx_slots(self) -> int:
        """Number of relative-offset slots: offsets -W..W -> slot = off + W."""
        return 2 * self.window + 1

    @property
    def neg_calls(self) -> int:
        """Philox calls per pair (each call yields two negatives)."""
        return (self.negatives + 1) // 2

    def to_dict(self):
        return asdict(self)


# ----------------------------------------------------------------------------
# randomness shared with the device
# ----------------------------------------------------------------------------

def subsample_mask(tokens: np.ndarray, keep_thresh: np.ndarray, seed: int, iteration: int,
                   raw_pos0: int) -> np.ndarray:
    """Boolean keep mask (component C5/K6).  Token at raw stream position p is
    kept iff ``philox(seed, SUBSAMPLE|iter, p).x <= keep_thresh[token]``."""
    pos = np.arange(raw_pos0, raw_pos0 + tokens.shape[0], dtype=np.uint64)
    r0, _, _, _ = philox.rand4(seed, philox.STREAM_SUBSAMPLE, pos, 0, iteration)
    return r0 <= keep_thresh[tokens]


def window_bounds(cfg: SGNSConfig, pos: np.ndarray, iteration: int) -> Tuple[np.ndarray, np.ndarray]:
    """Inclusive relative context offsets [lo, hi] per centre (C6/K7)."""
    r0, _, _, _ = philox.rand4(cfg.seed, philox.STREAM_WINDOW, pos, 0, iteration)
    b = philox.mulhi32(r0, np.uint32(cfg.window)).astype(np.int64)       # [0, W-1]
    if cfg.window_mode == "reference":
        return -b, b - 1                     # b left, b-1 right; empty when b == 0
    r = cfg.window - b                       # word2vec.c: radius in [1, W]
    return -r, r


def enumerate_pairs(cfg: SGNSConfig, tokens: np.ndarray, sent_id: np.ndarray, pos0: int,
                    iteration: int, lo: int = 0, hi: Optional[int] = None):
    """All (centre, context) pairs for centres ``lo <= i < hi`` of one step.

    Returns ``(ci, cj, slot)``: centre index, context index (both into the
    step's token array) and relative-offset slot, ordered by centre then slot.
    """
    t = tokens.shape[0]
    hi = t if hi is None else hi
    idx = np.arange(lo, hi, dtype=np.int64)
    pos = (np.uint64(pos0) + idx.astype(np.uint64))
    wlo, whi = window_bounds(cfg, pos, iteration)
    w = cfg.window
    offs = np.arange(-w, w + 1, dtype=np.int64)
    ci = np.repeat(idx, offs.shape[0])
    off = np.tile(offs, idx.shape[0])
    cj = ci + off
    valid = (off != 0) & (off >= np.repeat(wlo, offs.shape[0])) & (off <= np.repeat(whi, offs.shape[0]))
    valid &= (cj >= 0) & (cj < t)
    cjc = np.clip(cj, 0, max(t - 1, 0))
    if t > 0:
        valid &= sent_id[cjc] == sent_id[np.clip(ci, 0, t - 1)]
    return ci[valid], cj[valid], (off[valid] + w)


def draw_negatives(cfg: SGNSConfig, alias: AliasTable, pos: np.ndarray, slot: np.ndarray,
                   iteration: int) -> np.ndarray:
    """[P, n] negatives for pairs identified by (centre stream position, slot)."""
    p = pos.shape[0]
    n = cfg.negatives
    out = np.empty((p, n), dtype=np.int32)
    if cfg.neg_sharing == "centre":
        slot = np.zeros_like(slot)                      # every pair of a centre draws the same negatives
    for c in range(cfg.neg_calls):
        sub = (slot.astype(np.uint64) * np.uint64(cfg.neg_calls) + np.uint64(c))
        r0, r1, r2, r3 = philox.rand4(cfg.seed, philox.STREAM_NEG, pos, sub, iteration)
        out[:, 2 * c] = alias.sample(r0, r1)
        if 2 * c + 1 < n:
            out[:, 2 * c + 1] = alias.sample(r2, r3)
    return out


def tile_negatives(cfg: SGNSConfig, alias: AliasTable, pos0: int, tile_ids: np.ndarray,
                   iteration: int) -> np.ndarray:
    """[len(tile_ids), tile_negatives] shared negatives of the given tiles of a step (neg_sharing="tile").
    Tile j covers centres [j*T, (j+1)*T) of the step; its draws are keyed by the stream position of its first
    centre, so every shard regenerates them without traffic."""
    nn = cfg.tile_negatives
    pos = np.uint64(pos0) + tile_ids.astype(np.uint64) * np.uint64(cfg.tile_centres)
    out = np.empty((len(tile_ids), nn), dtype=np.int32)
    for c in range((nn + 1) // 2):
        r0, r1, r2, r3 = philox.rand4(cfg.seed, philox.STREAM_NEG, pos, np.full(len(pos), c, dtype=np.uint64),
                                      iteration)
        out[:, 2 * c] = alias.sample(r0, r1)
        if 2 * c + 1 < nn:
            out[:, 2 * c + 1] = alias.sample(r2, r3)
    return out


def tile_terms(cfg: SGNSConfig, ci: np.ndarray):
    """Active centres of a pair list: (step index of each active centre, its pair count m_i, its tile id)."""
    centres, m = np.unique(ci, return_counts=True)
    return centres, m, centres // cfg.tile_centres


# ----------------------------------------------------------------------------
# coefficient (sigmoid + learning rate)   -- component C3 / K3
# ----------------------------------------------------------------------------

def _exp_table() -> torch.Tensor:
    i = torch.arange(EXP_TABLE_SIZE, dtype=torch.float64)
    tmp = torch.exp((2.0 * i / EXP_TABLE_SIZE - 1.0) * MAX_EXP)
    return (tmp / (tmp + 1.0)).to(torch.float32)


_EXP_TABLE = None


def sigmoid_coeff(f: torch.Tensor, label: float, alpha: float, mode: str = "exact",
                  max_grad: float = 0.0) -> torch.Tensor:
    """``(label - sigmoid(f)) * alpha`` with the reference's +-6 hard clip."""
    global _EXP_TABLE
    if mode == "table":
        if _EXP_TABLE is None:
            _EXP_TABLE = _exp_table()
        tab = _EXP_TABLE.to(f.device)
        # (EXP_TABLE_SIZE / MAX_EXP / 2.0) with integer division first = 83.0
        ind = ((f + MAX_EXP) * 83.0).to(torch.int64).clamp_(0, EXP_TABLE_SIZE - 1)
        sig = tab[ind]
    else:
        sig = torch.sigmoid(f)
    g = label - sig
    g = torch.where(f > MAX_EXP, torch.full_like(g, label - 1.0), g)
    g = torch.where(f < -MAX_EXP, torch.full_like(g, label), g)
    g = g * alpha
    if max_grad > 0:
        g = g.clamp_(-max_grad, max_grad)
    return g


def sgns_loss(fplus: torch.Tensor, fminus: torch.Tensor, neg_mask: torch.Tensor) -> torch.Tensor:
    """Sum of -log sigma(f+) - sum log sigma(-f-) on the clipped dots."""
    fp = fplus.clamp(-MAX_EXP, MAX_EXP)
    fm = fminus.clamp(-MAX_EXP, MAX_EXP)
    lp = torch.nn.functional.softplus(-fp).sum()
    lm = (torch.nn.functional.softplus(fm) * neg_mask).sum()
    return lp + lm


# ----------------------------------------------------------------------------
# dense single-process oracle
# ----------------------------------------------------------------------------

@dataclass
class StepStats:
    pairs: int = 0
    loss: float = 0.0
    max_abs_dot: float = 0.0


def _scaled_index_add(mat: torch.Tensor, idx: torch.Tensor, upd: torch.Tensor, scale: Optional[torch.Tensor]):
    """``mat[idx] += scale[idx] * upd`` -- ``scale`` is the optional per-row update scale (hot-row damping)."""
    if scale is not None:
        upd = upd * scale[idx].to(upd.dtype)[:, None]
    mat.index_add_(0, idx, upd)


def sgns_minibatch_reference(syn0: torch.Tensor, syn1: torch.Tensor, cfg: SGNSConfig,
                             alias: AliasTable, tokens: np.ndarray, sent_id: np.ndarray,
                             pos0: int, iteration: int, alpha: float,
                             lo: int = 0, hi: Optional[int] = None,
                             row_scale0: Optional[torch.Tensor] = None,
                             row_scale1: Optional[torch.Tensor] = None, tile_neg_scale: float = 1.0,
                             tile_neg_weight: float = 1.0) -> StepStats:
    """One mini-batch (centres ``lo..hi`` of the step) applied in place.

    All dots use pre-update rows; all updates are summed (index_add), i.e. the
    semantics of one ``dotprod`` + ``adjust`` round trip (Appendix B).
    A negative equal to the pair's positive context is skipped (word2vec.c's
    ``if (target == word) continue``).
    """
    ci, cj, slot = enumerate_pairs(cfg, tokens, sent_id, pos0, iteration, lo, hi)
    stats = StepStats(pairs=int(ci.shape[0]))
    if ci.shape[0] == 0:
        return stats                      # zero-pair batches are a clean no-op (Q4)
    if cfg.neg_sharing == "tile":
        return _minibatch_tile_reference(syn0, syn1, cfg, alias, tokens, pos0, iteration, alpha, ci, cj, stats,
                                         row_scale0, row_scale1, tile_neg_scale, tile_neg_weight)
    pos = np.uint64(pos0) + ci.astype(np.uint64)
    neg = draw_negatives(cfg, alias, pos, slot, iteration)
    tok = tokens.astype(np.int64)
    w = torch.from_numpy(tok[ci])
    c = torch.from_numpy(tok[cj])
    ng = torch.from_numpy(neg.astype(np.int64))
    neg_mask = (ng != c[:, None]).to(syn0.dtype)
    u = syn0[w]                                   # [P, d]
    vc = syn1[c]                                  # [P, d]
    vn = syn1[ng]                                 # [P, n, d]
    fplus = (u * vc).sum(-1)
    fminus = torch.einsum("pd,pnd->pn", u, vn)
    gplus = sigmoid_coeff(fplus, 1.0, alpha, cfg.sigmoid_mode, cfg.max_grad)
    gminus = sigmoid_coeff(fminus, 0.0, alpha, cfg.sigmoid_mode, cfg.max_grad) * neg_mask
    stats.loss = float(sgns_loss(fplus, fminus, neg_mask))
    stats.max_abs_dot = float(max(fplus.abs().max(), fminus.abs().max()))
    du = gplus[:, None] * vc + torch.einsum("pn,pnd->pd", gminus, vn)
    dvc = gplus[:, None] * u
    dvn = gminus[:, :, None] * u[:, None, :]
    _scaled_index_add(syn0, w, du, row_scale0)
    _scaled_index_add(syn1, c, dvc, row_scale1)
    _scaled_index_add(syn1, ng.reshape(-1), dvn.reshape(-1, dvn.shape[-1]), row_scale1)
    return stats


def _minibatch_tile_reference(syn0, syn1, cfg, alias, tokens, pos0, iteration, alpha, ci, cj, stats,
                              row_scale0=None, row_scale1=None, tile_neg_scale: float = 1.0,
                              tile_neg_weight: float = 1.0) -> StepStats:
    """neg_sharing="tile": positives per pair, negatives per (active centre, shared negative of its tile) with
    weight m_i * n / N; every dot from pre-update rows, all updates summed."""
    tok = tokens.astype(np.int64)
    w = torch.from_numpy(tok[ci])
    c = torch.from_numpy(tok[cj])
    u = syn0[w]
    vc = syn1[c]
    fplus = (u * vc).sum(-1)
    gplus = sigmoid_coeff(fplus, 1.0, alpha, cfg.sigmoid_mode, cfg.max_grad)
    centres, m, tile = tile_terms(cfg, ci)
    tiles, inv = np.unique(tile, return_inverse=True)
    tneg = tile_negatives(cfg, alias, pos0, tiles, iteration)
    ng = torch.from_numpy(tneg[inv].astype(np.int64))                       # [A, N]
    wgt = torch.from_numpy(m.astype(np.float64) * cfg.negatives / cfg.tile_negatives).to(syn0.dtype)
    wa = torch.from_numpy(tok[centres])
    ua = syn0[wa]                                                           # [A, d]
    vn = syn1[ng]                                                           # [A, N, d]
    fminus = torch.einsum("ad,and->an", ua, vn)
    # tile_neg_weight scales the negative term of the UPDATES on both sides (EngineOptions.tile_neg_weight); the loss
    # keeps the full weight
    gminus = sigmoid_coeff(fminus, 0.0, alpha, cfg.sigmoid_mode, cfg.max_grad) * (wgt * float(tile_neg_weight))[:, None]
    stats.loss = float(sgns_loss(fplus, fminus, wgt[:, None].expand_as(fminus)))
    stats.max_abs_dot = float(max(fplus.abs().max(), fminus.abs().max()))
    du_pos = gplus[:, None] * vc
    dvc = gplus[:, None] * u
    du_neg = torch.einsum("an,and->ad", gminus, vn)
    # tile_neg_scale < 1 (engine.tile_neg_scale): the summed update of a shared negative row is one stale event of
    # mass sum_i m_i n / N (~90 unit updates for 128 centres, N = 32) and is capped like the hot rows
    dvn = gminus[:, :, None] * ua[:, None, :] * float(tile_neg_scale)
    _scaled_index_add(syn0, w, du_pos, row_scale0)
    _scaled_index_add(syn0, wa, du_neg, row_scale0)
    _scaled_index_add(syn1, c, dvc, row_scale1)
    _scaled_index_add(syn1, ng.reshape(-1), dvn.reshape(-1, dvn.shape[-1]), row_scale1)
    return stats


def init_embeddings(vocab_size: int, vector_size: int, seed: int, dtype=torch.float32,
                    scale_dim: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """word2vec.c initialisation: ``syn0 ~ U(-0.5, 0.5)/d``, ``syn1neg = 0``.

    Element (row r, column c) is a pure function of (seed, r, c) so that any
    column shard can initialise its own slice: ``philox(seed, INIT, r, c//4)``
    word ``c % 4``."""
    cols4 = (vector_size + 3) // 4
    rows = np.arange(vocab_size, dtype=np.uint64)[:, None]
    sub = np.arange(cols4, dtype=np.uint64)[None, :]
    r = philox.rand4(seed, philox.STREAM_INIT, rows, sub)
    u = np.stack([philox.u32_to_unit_float(x) for x in r], axis=-1).reshape(vocab_size, cols4 * 4)
    syn0 = (u[:, :vector_size] - np.float32(0.5)) / np.float32(scale_dim or vector_size)
    return (torch.from_numpy(np.ascontiguousarray(syn0)).to(dtype),
            torch.zeros(vocab_size, vector_size, dtype=dtype))


def learning_rate(lr: float, words_processed: int, total_words: int) -> float:
    """Closed form of the reference schedule with true global progress (Q5,
    MLLIB:405-410): ``lr * max(1e-4, 1 - processed / (iters*trainWords + 1))``."""
    return lr * max(1e-4, 1.0 - words_processed / float(total_words + 1))
