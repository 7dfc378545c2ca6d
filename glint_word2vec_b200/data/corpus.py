"""Sentence -> index mapping, chunking and step batching.

Reference components C4 (MLLIB:335-345: drop OOV words, split into chunks of
at most ``maxSentenceLength`` words) and the outer loop of C7 (MLLIB:401-419).
The reference batches at most ``batchSize`` (50) centres of ONE sentence per
RPC; the B200 engine packs many whole sentences into one device step and keeps
sentence boundaries as a per-token sentence id so windows never cross them
(SURVEY.md Q3).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Iterator, List, Sequence

import numpy as np

from .vocab import Vocabulary


def java_split(line: str, sep: str = " ") -> List[str]:
    """``String.split(" ")`` of the JVM: interior/leading empty tokens are kept,
    trailing empty tokens are removed (SURVEY.md Q9 -- this is how the empty
    string becomes a vocabulary word in the reference's test run, SPEC:85)."""
    if line == "":
        return [""]                 # Java: no match -> the input string itself
    parts = line.split(sep)
    while parts and parts[-1] == "":
        parts.pop()
    return parts


@dataclass
class EncodedCorpus:
    """Flat int32 token stream plus sentence offsets (CSR).

    The arrays may be ``np.memmap`` views of ``<prefix>.tokens.i32`` / ``<prefix>.offsets.i64`` (``open`` /
    ``save``): a corpus that does not fit in host memory is then streamed from disk step by step, and a shard-server
    group receives the PREFIX instead of the tokens (every rank maps the same files; the reference likewise never
    materialises the corpus on one node: RDD partitions, MLLIB:335-345)."""
    tokens: np.ndarray          # int32 [N]
    offsets: np.ndarray         # int64 [num_sentences + 1]
    prefix: str | None = None   # set when the arrays are disk-backed

    @classmethod
    def open(cls, prefix: str) -> "EncodedCorpus":
        import os
        tp, op = prefix + ".tokens.i32", prefix + ".offsets.i64"
        ntok = os.path.getsize(tp) // 4
        toks = np.memmap(tp, dtype=np.int32, mode="r", shape=(ntok,)) if ntok else np.zeros(0, np.int32)
        offs = np.memmap(op, dtype=np.int64, mode="r", shape=(os.path.getsize(op) // 8,))
        return cls(toks, offs, prefix)

    def save(self, prefix: str, chunk: int = 1 << 24) -> "EncodedCorpus":
        """Write the arrays to ``<prefix>.*`` in chunks and return the disk-backed twin."""
        with open(prefix + ".tokens.i32", "wb") as f:
            for lo in range(0, self.num_tokens, chunk):
                f.write(np.ascontiguousarray(self.tokens[lo:lo + chunk], dtype=np.int32).tobytes())
        with open(prefix + ".offsets.i64", "wb") as f:
            for lo in range(0, self.offsets.shape[0], chunk):
                f.write(np.ascontiguousarray(self.offsets[lo:lo + chunk], dtype=np.int64).tobytes())
        return EncodedCorpus.open(prefix)

    @property
    def num_tokens(self) -> int:
        return int(self.tokens.shape[0])

    @property
    def num_sentences(self) -> int:
        return int(self.offsets.shape[0] - 1)

    def sentence(self, i: int) -> np.ndarray:
        return self.tokens[self.offsets[i]:self.offsets[i + 1]]


def encode_corpus(sentences: Iterable[Sequence[str]], vocab: Vocabulary,
                  max_sentence_length: int = 1000, use_native: bool = True) -> EncodedCorpus:
    """words -> vocabulary indices, OOV dropped, chunked (MLLIB:335-343)."""
    if use_native:
        try:
            from ..ops import host as _host
            if _host.available() and isinstance(vocab.index, dict):
                return _host.encode_corpus(sentences, vocab, max_sentence_length)
        except Exception:  # pragma: no cover
            pass
    get = vocab.index.get
    toks: List[int] = []
    offs: List[int] = [0]
    for s in sentences:
        idx = [i for i in (get(w) for w in s) if i is not None]
        for lo in range(0, len(idx), max_sentence_length):
            chunk = idx[lo:lo + max_sentence_length]
            toks.extend(chunk)
            offs.append(len(toks))
    return EncodedCorpus(np.asarray(toks, dtype=np.int32), np.asarray(offs, dtype=np.int64))


def chunk_encoded(tokens: np.ndarray, offsets: np.ndarray, max_sentence_length: int) -> EncodedCorpus:
    """Re-chunk an already encoded corpus so that no sentence exceeds the limit; empty sentences are dropped
    (like ``encode_corpus`` drops sentences without in-vocabulary words)."""
    lens = np.diff(offsets)
    if lens.size == 0 or (lens.min() > 0 and lens.max() <= max_sentence_length):
        return EncodedCorpus(np.asarray(tokens, np.int32), np.asarray(offsets, np.int64))
    new_offs = [0]
    for a, b in zip(offsets[:-1], offsets[1:]):
        for lo in range(int(a), int(b), max_sentence_length):
            new_offs.append(min(lo + max_sentence_length, int(b)))
    return EncodedCorpus(np.asarray(tokens, np.int32), np.asarray(new_offs, np.int64))


def iter_text_file(path: str, tokenizer: str = "java") -> Iterator[List[str]]:
    """Sentences of a text file (one per line) as token lists; the pure-Python twin of the native loader."""
    with open(path, encoding="utf-8", errors="replace", newline="\n") as f:
        for line in f:
            if line.endswith("\n"):
                line = line[:-1]
            if line.endswith("\r"):
                line = line[:-1]
            yield java_split(line) if tokenizer == "java" else line.replace("\t", " ").replace("\r", " ").split()


def encode_text_file(path: str, vocab: Vocabulary, max_sentence_length: int = 1000, tokenizer: str = "java",
                     use_native: bool = True, num_threads: int = 0, out_prefix: str | None = None,
                     block_bytes: int = 256 << 20) -> EncodedCorpus:
    """Text file -> encoded corpus (OOV dropped, sentences chunked, MLLIB:335-343) without materialising Python
    token lists: ``csrc/host/textproc.cpp::encode_file`` mmaps the file and encodes on all cores.

    ``out_prefix``: stream the result to ``<out_prefix>.tokens.i32`` / ``.offsets.i64`` block by block
    (``encode_file_to``) and return a memory-mapped corpus -- peak host memory is one block, whatever the file size."""
    if tokenizer not in ("java", "whitespace"):
        raise ValueError(f"unknown tokenizer {tokenizer!r}")
    if out_prefix is not None:
        from ..ops import host as _host
        if use_native and _host.available():
            _host.encode_file_to(path, list(vocab.words), int(max_sentence_length), tokenizer == "java", num_threads,
                                 out_prefix, block_bytes)
            return EncodedCorpus.open(out_prefix)
        return encode_corpus(iter_text_file(path, tokenizer), vocab, max_sentence_length, use_native=False).save(out_prefix)
    if use_native:
        from ..ops import host as _host
        if _host.available():
            toks, offs = _host.encode_file(path, list(vocab.words), int(max_sentence_length), tokenizer == "java",
                                           num_threads)
            return EncodedCorpus(toks, offs)
    return encode_corpus(iter_text_file(path, tokenizer), vocab, max_sentence_length, use_native=False)


@dataclass
class StepBatch:
    """One device step: whole sentences, at most ``step_tokens`` tokens."""
    tokens: np.ndarray      # int32 [T]
    sent_id: np.ndarray     # int32 [T]  (monotone, relative to the step)
    raw_pos0: int           # position of tokens[0] in the (iteration-local) raw stream
    n_words: int            # == T (raw words consumed, for the LR schedule)


def iter_steps(corpus: EncodedCorpus, step_tokens: int) -> Iterator[StepBatch]:
    """Pack consecutive sentences into steps of <= ``step_tokens`` tokens.

    A sentence longer than ``step_tokens`` is split (windows at the cut are
    lost, exactly like the reference loses them at ``maxSentenceLength`` cuts).
    """
    offs = corpus.offsets
    ns = corpus.num_sentences
    s = 0
    while s < ns:
        start = int(offs[s])
        # number of whole sentences that fit
        e = int(np.searchsorted(offs, start + step_tokens, side="right")) - 1
        if e <= s:                       # a single sentence larger than the step
            end = min(int(offs[s + 1]), start + step_tokens)
            toks = np.ascontiguousarray(corpus.tokens[start:end])
            sid = np.zeros(end - start, dtype=np.int32)
            yield StepBatch(toks, sid, start, end - start)
            if end == int(offs[s + 1]):
                s += 1
            else:                        # leave the remainder as a shortened sentence
                offs = offs.copy()
                offs[s] = end
            continue
        end = int(offs[e])
        toks = np.ascontiguousarray(corpus.tokens[start:end])      # memmap slice -> one sequential read of the step
        lens = np.diff(offs[s:e + 1])
        sid = np.repeat(np.arange(e - s, dtype=np.int32), lens)
        yield StepBatch(toks, sid, start, end - start)
        s = e


def sentences_from_any(data, input_col: str | None = None) -> Iterable[Sequence[str]]:
    """Adapter: pandas DataFrame / pyarrow Table / dict-of-columns /
    list-of-lists / iterator -> iterable of token sequences.

    This replaces ``dataset.select($(inputCol)).rdd.map(_.getAs[Seq[String]](0))``
    (ML:286)."""
    try:
        import pandas as pd
        if isinstance(data, pd.DataFrame):
            if input_col is None or input_col not in data.columns:
                raise ValueError(f"input column {input_col!r} not found in DataFrame")
            return [list(x) if x is not None else [] for x in data[input_col].tolist()]
    except ImportError:  # pragma: no cover
        pass
    try:
        import pyarrow as pa
        if isinstance(data, pa.Table):
            return [list(x) if x is not None else [] for x in data.column(input_col).to_pylist()]
    except ImportError:  # pragma: no cover
        pass
    if isinstance(data, dict):
        return [list(x) for x in data[input_col]]
    return data
