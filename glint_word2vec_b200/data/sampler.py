"""Negative-sampling alias table, sub-sampling thresholds, synthetic Zipf streams.

Reference side: the Glint servers build a ``unigramTableSize``-entry table
(default 1e8 ints = 400 MB per server, ML:204-206) from ``cn^0.75`` and draw
``table[rand % size]`` from a per-request seed ([G]).  The
B200 design replaces the 400 MB table with a Vose alias table (8 bytes per
word, exact distribution, L2 resident for mid-size vocabularies) addressed by
Philox counters, so all ranks draw identical negatives with zero traffic.

All thresholds are *integers* (uint32) so that the numpy oracle and the CUDA
kernels take bit-identical decisions.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ..utils import philox

U32_MAX = 0xFFFFFFFF


@dataclass
class AliasTable:
    thresh: np.ndarray   # uint32 [V]  accept own bucket iff r1 < thresh
    alias: np.ndarray    # int32  [V]

    @property
    def size(self) -> int:
        return int(self.thresh.shape[0])

    def packed(self) -> np.ndarray:
        """[V, 2] int32 view {thresh bits, alias} as consumed by the kernels."""
        out = np.empty((self.size, 2), dtype=np.int32)
        out[:, 0] = self.thresh.view(np.int32)
        out[:, 1] = self.alias
        return out

    def sample(self, r0, r1) -> np.ndarray:
        """Map two uint32 randoms per draw onto a word index."""
        bucket = philox.mulhi32(r0, np.uint32(self.size)).astype(np.int64)
        take = np.asarray(r1, dtype=np.uint32) < self.thresh[bucket]
        return np.where(take, bucket, self.alias[bucket]).astype(np.int32)

    def probabilities(self) -> np.ndarray:
        """Exact distribution represented by the table (for chi-square tests)."""
        v = self.size
        p_own = self.thresh.astype(np.float64) / 2.0 ** 32
        p = p_own / v
        np.add.at(p, self.alias, (1.0 - p_own) / v)
        return p


def _vose_python(p: np.ndarray):
    v = p.shape[0]
    scaled = p * v
    prob = np.ones(v, dtype=np.float64)
    alias = np.arange(v, dtype=np.int64)
    small = [i for i in range(v) if scaled[i] < 1.0]
    large = [i for i in range(v) if scaled[i] >= 1.0]
    scaled = scaled.copy()
    while small and large:
        s = small.pop()
        l = large.pop()
        prob[s] = scaled[s]
        alias[s] = l
        scaled[l] = (scaled[l] + scaled[s]) - 1.0
        if scaled[l] < 1.0:
            small.append(l)
        else:
            large.append(l)
    # leftovers keep prob 1 / alias self
    return prob, alias


def build_alias(weights: np.ndarray, use_native: bool = True) -> AliasTable:
    """Vose alias table for an arbitrary non-negative weight vector."""
    w = np.asarray(weights, dtype=np.float64)
    if w.ndim != 1 or w.size == 0:
        raise ValueError("weights must be a non-empty vector")
    total = w.sum()
    if not np.isfinite(total) or total <= 0:
        raise ValueError("weights must have a positive finite sum")
    p = w / total
    prob = alias = None
    if use_native:
        try:
            from ..ops import host as _host
            if _host.available():
                prob, alias = _host.vose_alias(p)
        except Exception:  # pragma: no cover
            prob = None
    if prob is None:
        prob, alias = _vose_python(p)
    thresh = np.minimum(np.floor(prob * 2.0 ** 32), U32_MAX).astype(np.uint64).astype(np.uint32)
    alias = np.asarray(alias, dtype=np.int32)
    # prob == 1 buckets must alias to themselves so r1 == 0xFFFFFFFF stays exact
    full = prob >= 1.0
    alias[full] = np.arange(w.size, dtype=np.int32)[full]
    return AliasTable(thresh, alias)


def unigram_alias(counts: np.ndarray, power: float = 0.75, use_native: bool = True) -> AliasTable:
    """``cn^0.75`` noise distribution of word2vec (InitUnigramTable)."""
    return build_alias(np.asarray(counts, dtype=np.float64) ** power, use_native=use_native)


def unigram_table_counts(counts: np.ndarray, table_size: int, power: float = 0.75) -> np.ndarray:
    """How many of the ``table_size`` slots of word2vec.c's ``InitUnigramTable`` each word occupies (int64 [V]).

    The reference's servers draw negatives as ``table[rand % unigramTableSize]`` from a table filled like
    ``InitUnigramTable`` (MLLIB:239-244, ML:204-209, [G]): slot a belongs to word i, and i advances by at most one per
    slot once ``a / size`` exceeds the cumulative ``cn^0.75`` mass.  The slot counts are that quantised distribution
    exactly; ``sampler="table"`` feeds them to the alias sampler, which then draws from the same distribution as the
    400 MB table without building it."""
    cn = np.asarray(counts, dtype=np.float64) ** power
    v = cn.shape[0]
    cum = np.cumsum(cn / cn.sum())
    idx = np.arange(v, dtype=np.int64)
    first_after = np.floor(cum * table_size).astype(np.int64) + 1        # first slot a with a / size > cum[i]
    last = np.maximum.accumulate(first_after - idx) + idx                 # i advances at most once per slot
    last = np.minimum(last, table_size - 1)
    last[-1] = table_size - 1                                             # `if (i >= vocab_size) i = vocab_size - 1`
    prev = np.concatenate([[-1], last[:-1]])
    return np.maximum(last - prev, 0)


def unigram_table_alias(counts: np.ndarray, table_size: int, power: float = 0.75, use_native: bool = True) -> AliasTable:
    """Alias table over the slot counts of the reference's unigram table (``sampler="table"`` parity mode)."""
    return build_alias(unigram_table_counts(counts, table_size, power).astype(np.float64), use_native=use_native)


def keep_thresholds(counts: np.ndarray, subsample_ratio: float, mode: str = "word2vec") -> np.ndarray:
    """uint32 threshold per word: a token is kept iff ``r <= thresh``.

    ``mode="word2vec"``: the *intended* formula of MLLIB:375-377,
    ``keep = min(1, (sqrt(f/t) + 1) * t / f)`` with ``f = cn/trainWords``.
    ``mode="reference"``: reproduces the reference bug (integer division makes
    ``f == 0`` so every word is kept).
    """
    counts = np.asarray(counts, dtype=np.float64)
    if mode == "reference" or subsample_ratio <= 0:
        return np.full(counts.shape[0], U32_MAX, dtype=np.uint32)
    if mode != "word2vec":
        raise ValueError(f"unknown subsample mode {mode!r}")
    f = counts / counts.sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        keep = (np.sqrt(f / subsample_ratio) + 1.0) * (subsample_ratio / f)
    keep = np.where(f > 0, keep, 1.0)
    keep = np.clip(keep, 0.0, 1.0)
    return np.minimum(np.floor(keep * 2.0 ** 32), U32_MAX).astype(np.uint64).astype(np.uint32)


def zipf_counts(vocab_size: int, total_words: int, exponent: float = 1.0) -> np.ndarray:
    """Expected word counts of a Zipf(``exponent``) vocabulary, count >= 1."""
    ranks = np.arange(1, vocab_size + 1, dtype=np.float64)
    w = ranks ** (-exponent)
    w *= total_words / w.sum()
    return np.maximum(1, np.round(w)).astype(np.int64)


def zipf_tokens(alias: AliasTable, n: int, seed: int, pos0: int = 0) -> np.ndarray:
    """Deterministic Zipf token stream: token at stream
    position p is ``alias.sample(philox(seed, ZIPF, p))`` -- identical on all
    ranks, identical between numpy and ``zipf_stream`` on the device."""
    pos = np.arange(pos0, pos0 + n, dtype=np.uint64)
    r0, r1, _, _ = philox.rand4(seed, philox.STREAM_ZIPF, pos)
    return alias.sample(r0, r1)
