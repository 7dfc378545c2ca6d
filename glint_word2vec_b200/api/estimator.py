"""``ServerSideGlintWord2Vec`` estimator (components C1, C14, C17).

Reference: Spark ML estimator ML:228-317 delegating to the MLlib trainer
MLLIB:65-449; Python wrapper PY:38-305.  ``fit`` = vocabulary -> sentence
encoding -> shard-group bootstrap (integrated or separate, MLLIB:351-362) ->
training -> model.
"""
from __future__ import annotations

import logging
import os
from typing import Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from ..data.corpus import EncodedCorpus, chunk_encoded, encode_corpus, sentences_from_any
from ..data.vocab import Vocabulary, build_vocab, vocab_from_counts
from ..models.sgns import SGNSConfig
from ..parallel import cluster as _cluster
from . import frames
from .model import ServerSideGlintWord2VecModel
from .params import ServerSideGlintWord2VecBase

log = logging.getLogger("glint_word2vec_b200")

JAVA_ESTIMATOR_CLASS = "org.apache.spark.ml.feature.ServerSideGlintWord2Vec"

_ENGINE_KEYS = ("step_tokens", "subsample_mode", "transport", "kernel",
                "store_syn1", "max_hot_updates", "window_mode", "sigmoid_mode", "max_grad", "neg_sharing", "tile_centres",
                "tile_negatives", "device", "hot_row_cap", "sampler", "tile_neg_weight")


# keys of parameterServerConfig that configure the model spec (SGNSConfig) or the placement, not EngineOptions
_CONFIG_KEYS = ("window_mode", "sigmoid_mode", "max_grad", "neg_sharing", "tile_centres", "tile_negatives", "device")


def engine_options_from_params(p: ServerSideGlintWord2VecBase) -> dict:
    """Engine options = ``parameterServerConfig`` pass-through + the ML params
    that the Glint servers receive through ``Word2VecArguments`` (MLLIB:351)."""
    cfg = p.getParameterServerConfig()
    opts = {k: v for k, v in cfg.items() if k in _ENGINE_KEYS}
    opts["batch_size"] = p.getBatchSize()
    opts["subsample_ratio"] = p.getSubsampleRatio()
    # numPartitions = asynchronous workers of the reference (MLLIB:122-126,345,392): the staleness knob of the engine;
    # unigramTableSize is honoured by sampler="table" (MLLIB:239-244)
    opts["num_partitions"] = p.getNumPartitions()
    opts["unigram_table_size"] = p.getUnigramTableSize()
    return opts


def _spmd_world() -> int:
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def _pick_device(opts: dict):
    want = opts.get("device", "auto")
    if want == "cpu" or not torch.cuda.is_available():
        return torch.device("cpu")
    return torch.device("cuda", torch.cuda.current_device())


def open_handle_for_fit(cfg: SGNSConfig, counts, host: str, num_servers: int, opts: dict):
    """PS bootstrap (C9): separate cluster if a host is given, SPMD if this
    process is one rank of a torchrun job, in-process for one shard, else spawn
    an integrated shard-server group."""
    engine_opts = {k: v for k, v in opts.items() if k not in _CONFIG_KEYS}
    if host:
        h = _cluster.connect_separate(host)
        return h.create(cfg, engine_opts, counts)
    if _spmd_world() > 1:
        if num_servers != _spmd_world():
            log.info("SPMD job: using world size %d as the shard count (numParameterServers=%d ignored)",
                     _spmd_world(), num_servers)
        return _cluster.InProcessHandle.create(cfg, engine_opts, counts, device=_pick_device(opts))
    shards, dev = _cluster.resolve_integrated_shards(num_servers)
    if opts.get("device") == "cpu":
        dev, shards = "cpu", max(1, num_servers)
    if shards == 1:
        return _cluster.InProcessHandle.create(cfg, engine_opts, counts, device=_pick_device(opts))
    h = _cluster.spawn_integrated(shards, dev, None)
    return h.create(cfg, engine_opts, counts)


def open_handle_for_load(path: str, host: str, num_servers: int, opts: dict):
    engine_opts = {k: v for k, v in opts.items() if k not in _CONFIG_KEYS}
    if host:
        h = _cluster.connect_separate(host)
        h.load(path, engine_opts)
        return h
    if _spmd_world() > 1:
        return _cluster.InProcessHandle.load(path, engine_opts, device=_pick_device(opts))
    shards, dev = _cluster.resolve_integrated_shards(num_servers)
    if opts.get("device") == "cpu":
        dev, shards = "cpu", max(1, num_servers)
    if shards == 1:
        return _cluster.InProcessHandle.load(path, engine_opts, device=_pick_device(opts))
    h = _cluster.spawn_integrated(shards, dev, None)
    h.load(path, engine_opts)
    return h


class ServerSideGlintWord2Vec(ServerSideGlintWord2VecBase):
    """Spark-ML style estimator.  Keyword-only constructor like PY:138-155."""

    _uid_prefix = "gw2v"            # ML:231

    def __init__(self, uid: Optional[str] = None, **kwargs):
        super().__init__(uid)
        self._declare_w2v_params()
        self.setParams(**kwargs)

    def setParams(self, **kwargs):
        """PY:157-170 (plus ``parameterServerConfig``, which the reference's Python wrapper lacks, Q10)."""
        for k in kwargs:
            if not self.hasParam(k):
                raise TypeError(f"setParams() got an unexpected keyword argument {k!r}")
        return self._set(**kwargs)

    # setters (ML:234-282; MLlib aliases MLLIB:112,133)
    def setInputCol(self, v): return self.set("inputCol", v)
    def setOutputCol(self, v): return self.set("outputCol", v)
    def setVectorSize(self, v): return self.set("vectorSize", v)
    def setWindowSize(self, v): return self.set("windowSize", v)
    def setStepSize(self, v): return self.set("stepSize", v)
    def setNumPartitions(self, v): return self.set("numPartitions", v)
    def setMaxIter(self, v): return self.set("maxIter", v)
    def setSeed(self, v): return self.set("seed", v)
    def setMinCount(self, v): return self.set("minCount", v)
    def setMaxSentenceLength(self, v): return self.set("maxSentenceLength", v)
    def setBatchSize(self, v): return self.set("batchSize", v)
    def setN(self, v): return self.set("n", v)
    def setSubsampleRatio(self, v): return self.set("subsampleRatio", v)
    def setNumParameterServers(self, v): return self.set("numParameterServers", v)
    def setParameterServerHost(self, v): return self.set("parameterServerHost", v)
    def setParameterServerConfig(self, v): return self.set("parameterServerConfig", v)
    def setUnigramTableSize(self, v): return self.set("unigramTableSize", v)

    # ------------------------------------------------------------------ fit
    def _validate_for_fit(self):
        """One validation pass in ``fit`` instead of the reference's
        setter-order-dependent checks (Q7).  The Akka payload guard
        ``batchSize * n * window <= 10000`` (MLLIB:154-155) has no counterpart:
        there is no message-size limit on NVLink."""
        for name, cond in (("batchSize", lambda v: v > 0), ("n", lambda v: v > 0),
                           ("subsampleRatio", lambda v: v >= 0), ("numParameterServers", lambda v: v > 0),
                           ("unigramTableSize", lambda v: v > 0)):
            v = self.getOrDefault(name)
            if not cond(v):
                raise ValueError(f"{name} must be positive but got {v}")

    def fit(self, dataset, params: Optional[dict] = None) -> ServerSideGlintWord2VecModel:
        """``dataset``: DataFrame/Table/dict with ``inputCol`` of token lists, or
        any iterable of token sequences (the MLlib ``fit(RDD[Iterable[String]])``)."""
        if params:
            return self.copy(params).fit(dataset)
        self._validate_for_fit()
        if frames.column_names(dataset) is not None:
            self._validate_input(dataset)
            sentences = sentences_from_any(dataset, self.getInputCol())
        else:
            sentences = dataset if isinstance(dataset, (list, tuple)) else list(dataset)
        vocab = build_vocab(sentences, self.getMinCount())
        log.info("vocabSize = %d, trainWordsCount = %d", vocab.size, vocab.train_words)   # MLLIB:278
        corpus = encode_corpus(sentences, vocab, self.getMaxSentenceLength())
        return self._fit_encoded(vocab, corpus)

    def fitTextFile(self, path: str, tokenizer: str = "java") -> ServerSideGlintWord2VecModel:
        """Train on a text file with one sentence per line.  Vocabulary counting and encoding run in the native
        host library (mmap + all cores) without ever building Python token lists - the loader for corpora of
        the reference's scale, which it reads as Spark RDD partitions (MLLIB:258-279,335-345).
        ``tokenizer="java"`` = the spec's ``line.split(" ")``; ``"whitespace"`` = split on runs of blanks."""
        from ..data.corpus import encode_text_file
        from ..data.vocab import build_vocab_from_file
        self._validate_for_fit()
        vocab = build_vocab_from_file(path, self.getMinCount(), tokenizer)
        log.info("vocabSize = %d, trainWordsCount = %d", vocab.size, vocab.train_words)   # MLLIB:278
        # Large files are never held in host memory: the native encoder streams ``<cache>/corpus.tokens.i32`` block by
        # block, training reads it through a memory map, and shard servers receive the prefix, not the tokens.
        # parameterServerConfig: corpus_cache_dir (kept after the fit) / stream_threshold_bytes (default 1 GiB).
        pcfg = self.getParameterServerConfig()
        cache = pcfg.get("corpus_cache_dir")
        tmp = None
        if cache is None and os.path.getsize(path) >= int(pcfg.get("stream_threshold_bytes", 1 << 30)):
            import tempfile
            cache = tmp = tempfile.mkdtemp(prefix="gw2v-corpus-", dir=pcfg.get("scratch_dir"))
        if cache is None:
            return self._fit_encoded(vocab, encode_text_file(path, vocab, self.getMaxSentenceLength(), tokenizer))
        os.makedirs(cache, exist_ok=True)
        try:
            corpus = encode_text_file(path, vocab, self.getMaxSentenceLength(), tokenizer,
                                      out_prefix=os.path.join(cache, "corpus"))
            return self._fit_encoded(vocab, corpus)
        finally:
            if tmp is not None:
                import shutil
                shutil.rmtree(tmp, ignore_errors=True)

    def fitEncoded(self, tokens: np.ndarray, offsets: np.ndarray, counts: np.ndarray,
                   words: Optional[Sequence[str]] = None) -> ServerSideGlintWord2VecModel:
        """Train on an already index-encoded corpus (token ids must already be
        frequency ranks).  Used for synthetic/very large corpora where the
        string pipeline is pointless."""
        self._validate_for_fit()
        vocab = vocab_from_counts(counts, words)
        tokens = np.asarray(tokens)
        if tokens.size and (int(tokens.min()) < 0 or int(tokens.max()) >= vocab.size):
            raise ValueError(f"token ids must lie in [0, {vocab.size})")     # the kernels do not bounds-check
        corpus = chunk_encoded(np.asarray(tokens, np.int32), np.asarray(offsets, np.int64),
                               self.getMaxSentenceLength())
        return self._fit_encoded(vocab, corpus)

    def _fit_encoded(self, vocab: Vocabulary, corpus: EncodedCorpus) -> ServerSideGlintWord2VecModel:
        pcfg = self.getParameterServerConfig()
        cfg = SGNSConfig(vocab_size=vocab.size, vector_size=self.getVectorSize(), window=self.getWindowSize(),
                         negatives=self.getN(), seed=self.getSeed(),
                         window_mode=pcfg.get("window_mode", "reference"),
                         sigmoid_mode=pcfg.get("sigmoid_mode", "exact"),
                         max_grad=float(pcfg.get("max_grad", 0.0)),
                         neg_sharing=pcfg.get("neg_sharing", "pair"),
                         tile_centres=int(pcfg.get("tile_centres", 128)),
                         tile_negatives=int(pcfg.get("tile_negatives", 64)))
        opts = engine_options_from_params(self)
        handle = open_handle_for_fit(cfg, vocab.counts, self.getParameterServerHost(),
                                     self.getNumParameterServers(), opts)
        try:
            train_opts = {k: pcfg[k] for k in ("checkpoint_dir", "checkpoint_every_steps", "resume") if k in pcfg}
            report = handle.fit(corpus, self.getStepSize(), self.getMaxIter(), vocab.train_words,
                                pcfg.get("metrics_path"), train_opts)
        except Exception:
            handle.destroy()
            handle.terminate(False)
            raise
        model = ServerSideGlintWord2VecModel(words=vocab.words, handle=handle,
                                             word_index=vocab.index)
        self._copyValues(model)
        # the model remembers which server group it lives on (ML:516 reads it back on load)
        model.set("parameterServerHost", handle.host or "")
        model.setParent(self)
        model.trainingReport = report
        return model

    # ------------------------------------------------------------------ persistence (DefaultParamsWritable)
    def save(self, path: str):
        if os.path.exists(path):
            raise IOError(f"Path {path} already exists. To overwrite it, please use write.overwrite().save(path).")
        self._save_metadata(path, JAVA_ESTIMATOR_CLASS)

    def write(self):
        est = self

        class _W:
            _ow = False

            def overwrite(self):
                self._ow = True
                return self

            def save(self, path):
                if os.path.exists(path) and self._ow:
                    import shutil
                    shutil.rmtree(path)
                est.save(path)
        return _W()

    @classmethod
    def load(cls, path: str) -> "ServerSideGlintWord2Vec":
        meta = cls._load_metadata(path, JAVA_ESTIMATOR_CLASS)
        est = cls(uid=meta["uid"])
        est._get_and_set_params(meta)
        return est

    @classmethod
    def read(cls):
        class _R:
            def load(self, path):
                return cls.load(path)
        return _R()

    def copy(self, extra=None):
        that = ServerSideGlintWord2Vec(self.uid)
        self._copyValues(that, extra)
        return that
