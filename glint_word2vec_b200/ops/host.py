"""Loader for the native host library (``csrc/host/textproc.cpp``).

Vocabulary counting, corpus encoding and Vose alias construction are the
reference's CPU-heavy cold-path pieces (Spark shuffle ``reduceByKey`` at
MLLIB:262, the 1e8-entry unigram table of the Glint servers [G]); here they are
C++ behind a pybind11 module built in-tree by ``build_ext.py``.  Pure-Python
fallbacks exist in ``data/`` so the package works before the build.
"""
from __future__ import annotations

import importlib
from typing import Iterable, Sequence

import numpy as np

_mod = None
_tried = False


def _load():
    global _mod, _tried
    if not _tried:
        _tried = True
        try:
            _mod = importlib.import_module("glint_word2vec_b200._host")
        except Exception:
            _mod = None
    return _mod


def available() -> bool:
    return _load() is not None


def count_words(sentences: Iterable[Sequence[str]]):
    m = _load()
    return m.count_words(sentences)


def encode_corpus(sentences, vocab, max_sentence_length: int):
    from ..data.corpus import EncodedCorpus
    m = _load()
    toks, offs = m.encode_corpus(sentences, vocab.index, int(max_sentence_length))
    return EncodedCorpus(np.asarray(toks, dtype=np.int32), np.asarray(offs, dtype=np.int64))


def vose_alias(p: np.ndarray):
    m = _load()
    prob, alias = m.vose_alias(np.ascontiguousarray(p, dtype=np.float64))
    return np.asarray(prob), np.asarray(alias)


def count_words_file(path: str, java_mode: bool = True, num_threads: int = 0):
    """(words, counts int64) of a text file, one sentence per line."""
    m = _load()
    words, counts = m.count_words_file(str(path), bool(java_mode), int(num_threads))
    return words, np.asarray(counts, dtype=np.int64)


def encode_file(path: str, words, max_sentence_length: int, java_mode: bool = True, num_threads: int = 0):
    """(tokens int32, offsets int64) of a text file against the vocabulary ``words`` (index = position)."""
    m = _load()
    toks, offs = m.encode_file(str(path), list(words), int(max_sentence_length), bool(java_mode), int(num_threads))
    return np.asarray(toks, dtype=np.int32), np.asarray(offs, dtype=np.int64)


def encode_file_to(path: str, words, max_sentence_length: int, java_mode: bool, num_threads: int, out_prefix: str,
                   block_bytes: int = 256 << 20):
    """Streaming ``encode_file``: appends to ``<out_prefix>.tokens.i32`` / ``.offsets.i64`` block by block; returns
    (tokens, sentences)."""
    m = _load()
    ntok, nsent = m.encode_file_to(str(path), list(words), int(max_sentence_length), bool(java_mode), int(num_threads),
                                   str(out_prefix), int(block_bytes))
    return int(ntok), int(nsent)
