"""Vocabulary builder (reference component C2, ``learnVocab`` MLLIB:258-279).

word count -> ``minCount`` filter -> sort by count descending -> row index =
rank.  The reference leaves the order of equal counts unspecified (it is
whatever Spark's ``collect`` returns, MLLIB:265-266); we make it deterministic:
ties are ordered by the word's UTF-8 bytes.

The counting itself is done by the native host library
(``csrc/host/textproc.cpp``: sharded open-addressing hash count in C++) when
it is built; the pure-Python ``collections.Counter`` path is the fallback and
the oracle the native path is tested against.
"""
from __future__ import annotations

from collections import Counter
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Sequence

import numpy as np


@dataclass
class Vocabulary:
    words: List[str]
    counts: np.ndarray                      # int64 [V]
    index: Dict[str, int] = field(default_factory=dict, repr=False)

    def __post_init__(self):
        self.counts = np.asarray(self.counts, dtype=np.int64)
        if not self.index:
            self.index = {w: i for i, w in enumerate(self.words)}

    @property
    def size(self) -> int:
        return len(self.words)

    @property
    def train_words(self) -> int:
        """``trainWordsCount`` (MLLIB:274-276)."""
        return int(self.counts.sum())

    def __len__(self):
        return len(self.words)

    def __contains__(self, w):
        return w in self.index


def count_words_python(sentences: Iterable[Sequence[str]]) -> Counter:
    c: Counter = Counter()
    for s in sentences:
        c.update(s)
    return c


def build_vocab(sentences: Iterable[Sequence[str]], min_count: int = 5,
                use_native: bool = True) -> Vocabulary:
    """Build the vocabulary exactly like ``learnVocab``.

    Raises ``ValueError`` when no word survives ``min_count`` (the reference's
    ``require(vocabSize > 0, ...)`` at MLLIB:269-270).
    """
    counts = None
    if use_native:
        try:
            from ..ops import host as _host
            if _host.available():
                counts = _host.count_words(sentences)
        except Exception:  # pragma: no cover - native lib optional
            counts = None
    if counts is None:
        counts = count_words_python(sentences)
    items = [(w, c) for w, c in counts.items() if c >= min_count]
    if not items:
        raise ValueError(
            "The vocabulary size should be > 0. You may need to check the setting of "
            "minCount, which could be large enough to remove all your words in sentences.")
    items.sort(key=lambda wc: (-wc[1], wc[0].encode("utf-8")))
    words = [w for w, _ in items]
    cn = np.fromiter((c for _, c in items), dtype=np.int64, count=len(items))
    return Vocabulary(words, cn)


def _vocab_from_counts_dict(counts, min_count: int) -> Vocabulary:
    items = [(w, c) for w, c in counts if c >= min_count]
    if not items:
        raise ValueError(
            "The vocabulary size should be > 0. You may need to check the setting of "
            "minCount, which could be large enough to remove all your words in sentences.")
    items.sort(key=lambda wc: (-wc[1], wc[0].encode("utf-8")))
    words = [w for w, _ in items]
    cn = np.fromiter((c for _, c in items), dtype=np.int64, count=len(items))
    return Vocabulary(words, cn)


def build_vocab_from_file(path: str, min_count: int = 5, tokenizer: str = "java", use_native: bool = True,
                          num_threads: int = 0) -> Vocabulary:
    """``learnVocab`` straight from a text file with one sentence per line.

    ``tokenizer="java"`` splits like the reference's spec does (``line.split(" ")`` on the JVM: single spaces,
    interior empty tokens kept, Q9); ``"whitespace"`` splits on runs of blanks.  The native path mmaps the file
    and counts on all cores (``csrc/host/textproc.cpp::count_words_file``)."""
    if tokenizer not in ("java", "whitespace"):
        raise ValueError(f"unknown tokenizer {tokenizer!r}")
    if use_native:
        from ..ops import host as _host
        if _host.available():
            words, counts = _host.count_words_file(path, tokenizer == "java", num_threads)
            return _vocab_from_counts_dict(zip(words, counts.tolist()), min_count)
    from .corpus import iter_text_file
    c = count_words_python(iter_text_file(path, tokenizer))
    return _vocab_from_counts_dict(c.items(), min_count)


def vocab_from_counts(counts: np.ndarray, words: Sequence[str] | None = None) -> Vocabulary:
    """Vocabulary for synthetic corpora: word i is ``"w<i>"`` unless given."""
    counts = np.asarray(counts, dtype=np.int64)
    if words is None:
        words = _SyntheticWords(len(counts))
        v = Vocabulary.__new__(Vocabulary)
        v.words = words
        v.counts = counts
        v.index = _SyntheticIndex(len(counts))
        return v
    return Vocabulary(list(words), counts)


class _SyntheticWords(Sequence):
    """Lazy ``["w0", "w1", ...]`` so a 80M-word synthetic vocab costs no RAM."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [f"w{j}" for j in range(*i.indices(self.n))]
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        return f"w{i}"


class _SyntheticIndex:
    def __init__(self, n):
        self.n = n

    def get(self, w, default=None):
        try:
            return self[w]
        except KeyError:
            return default

    def __getitem__(self, w):
        if isinstance(w, str) and w[:1] == "w" and w[1:].isdigit():
            i = int(w[1:])
            if i < self.n:
                return i
        raise KeyError(w)

    def __contains__(self, w):
        return self.get(w) is not None

    def __len__(self):
        return self.n
