"""``ServerSideGlintWord2VecModel``: the trained, still-distributed model.

Reference components C10/C11/C12 (MLlib model, MLLIB:460-726), C15/C16 (ML model
+ persistence, ML:322-600) and C18 (Python wrapper, PY:311-383).  The matrices
stay sharded on the GPUs (the "parameter servers") until ``stop()``; every
method below is a batched request to the shard group.

On-disk layout written by ``save``::

    <path>/metadata/part-00000 + _SUCCESS    Spark DefaultParamsWriter JSON (same keys, same class name)
    <path>/words/part-00000 + _SUCCESS       one word per line, line k = row k (empty word preserved, Q9)
    <path>/matrix/...                        column shards (models/matrix_io.py)
"""
from __future__ import annotations

import os
import shutil
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np

from ..parallel import cluster as _cluster
from . import frames
from .params import ServerSideGlintWord2VecBase

JAVA_MODEL_CLASS = "org.apache.spark.ml.feature.ServerSideGlintWord2VecModel"
FORMAT_VERSION = "1.0"          # MLLIB:488
PULL_BATCH = 10000              # rows per request in the reference (MLLIB:531, ML:449)


class _Writer:
    """``MLWriter``: ``model.write().overwrite().save(path)``."""

    def __init__(self, instance):
        self._instance = instance
        self._overwrite = False

    def overwrite(self):
        self._overwrite = True
        return self

    def save(self, path: str):
        """SPMD (every rank of a torchrun job calls this): the writer rank decides and removes, the others learn the
        decision through a broadcast and wait at a barrier -- a rank that raced ahead could otherwise see the path
        half-removed, or write its shard into a directory that rank 0 is still deleting."""
        inst = self._instance
        comm = getattr(getattr(inst._handle, "engine", None), "comm", None)
        spmd = comm is not None and comm.world > 1
        exists = os.path.exists(path) if inst._is_writer_rank() else None
        if spmd:
            exists = comm.broadcast_object(exists, 0)
        if exists:
            if not self._overwrite:
                raise IOError(f"Path {path} already exists. To overwrite it, please use write.overwrite().save(path).")
            if inst._is_writer_rank():
                shutil.rmtree(path)
            if spmd:
                comm.barrier()
        inst._save_impl(path)


class ServerSideGlintWord2VecModel(ServerSideGlintWord2VecBase):
    _uid_prefix = "gw2v"

    def __init__(self, uid: Optional[str] = None, words: Optional[Sequence[str]] = None,
                 handle: Optional[_cluster.MatrixHandle] = None, word_index=None):
        super().__init__(uid)
        self._declare_w2v_params()
        self._words = words if words is not None else []
        self._index = word_index if word_index is not None else {w: i for i, w in enumerate(self._words)}
        self._handle = handle
        self._parent = None
        self.trainingReport: Optional[dict] = None

    # ------------------------------------------------------------- metadata
    @property
    def numWords(self) -> int:
        """ML:331 / MLLIB:468."""
        return len(self._words)

    @property
    def _vsize(self) -> int:
        """``matrix.cols`` (MLLIB:473); ``model.vectorSize`` itself is the Param, as in Spark."""
        return int(self._handle.cols) if self._handle is not None else self.getVectorSize()

    @property
    def parent(self):
        return self._parent

    def setParent(self, est):
        self._parent = est
        return self

    def setInputCol(self, value): return self.set("inputCol", value)      # ML:423
    def setOutputCol(self, value): return self.set("outputCol", value)    # ML:426

    def _require_handle(self):
        if self._handle is None:
            raise RuntimeError("model has been stopped (matrix destroyed)")
        return self._handle

    def _is_writer_rank(self) -> bool:
        h = self._handle
        eng = getattr(h, "engine", None)
        return eng is None or eng.comm.rank == 0

    # ------------------------------------------------------------- lookups
    def wordIndex(self, word: str) -> int:
        try:
            return self._index[word]
        except KeyError:
            raise KeyError(f"{word} not in vocabulary") from None

    def transformWord(self, word: str) -> np.ndarray:
        """MLlib ``transform(word)``: raises on OOV (MLLIB:511-519, Q11)."""
        return self._require_handle().pull([self.wordIndex(word)])[0].astype(np.float64)

    def transformWords(self, words: Iterable[str]) -> Iterator[np.ndarray]:
        """MLlib ``transform(Iterator[String])`` in 10 000-word batches (MLLIB:529-543)."""
        h = self._require_handle()
        batch: List[int] = []
        for w in words:
            batch.append(self.wordIndex(w))
            if len(batch) >= PULL_BATCH:
                for v in h.pull(batch):
                    yield v.astype(np.float64)
                batch = []
        if batch:
            for v in h.pull(batch):
                yield v.astype(np.float64)

    def getVectors(self):
        """DataFrame(word, vector) of all word vectors (ML:342-364)."""
        h = self._require_handle()
        n = self.numWords
        vecs = []
        step = 1 << 18
        for lo in range(0, n, step):
            vecs.append(h.pull(np.arange(lo, min(n, lo + step), dtype=np.int64)))
        mat = np.concatenate(vecs, 0) if vecs else np.zeros((0, self._vsize), np.float32)
        return frames.make_frame({"word": list(self._words[i] for i in range(n)),
                                  "vector": [mat[i].astype(np.float64) for i in range(n)]})

    def getVectorsMap(self) -> dict:
        """MLlib ``getVectors: Map[String, Array[Float]]`` (MLLIB:638-641)."""
        h = self._require_handle()
        mat = h.pull(np.arange(self.numWords, dtype=np.int64))
        return {self._words[i]: mat[i] for i in range(self.numWords)}

    # ------------------------------------------------------------- synonyms
    def findSynonymsArray(self, word_or_vec: Union[str, Sequence[float], np.ndarray], num: int) -> List[Tuple[str, float]]:
        """Top ``num`` cosine neighbours (ML:405-420 -> MLLIB:554-630).

        A word query excludes the word itself; a vector query does not."""
        return self.findSynonymsArrayBatch([word_or_vec], num)[0]

    def findSynonymsArrayBatch(self, queries: Sequence[Union[str, Sequence[float], np.ndarray]],
                               num: int) -> List[List[Tuple[str, float]]]:
        """Batched variant: one sweep over the sharded matrix serves all queries."""
        if num <= 0:
            raise ValueError("Number of similar words should > 0")         # MLLIB:587
        h = self._require_handle()
        word_opts: List[Optional[str]] = []
        word_rows = [self.wordIndex(q) for q in queries if isinstance(q, str)]
        pulled = iter(h.pull(word_rows)) if word_rows else iter(())
        vecs = []
        for q in queries:
            if isinstance(q, str):
                word_opts.append(q)
                vecs.append(np.asarray(next(pulled), dtype=np.float32))
            else:
                v = np.asarray(q, dtype=np.float32).reshape(-1)
                if v.shape[0] != self._vsize:
                    raise ValueError(f"query vector has length {v.shape[0]}, expected {self._vsize}")
                word_opts.append(None)
                vecs.append(v)
        k = min(num + 1, self.numWords)
        idx, sim = h.top_k(np.stack(vecs), k)
        out = []
        for qi, wopt in enumerate(word_opts):
            res = [(self._words[int(i)], float(s)) for i, s in zip(idx[qi], sim[qi])]
            if wopt is not None:
                res = [r for r in res if r[0] != wopt]
            out.append(res[:num])
        return out

    def findSynonyms(self, word_or_vec, num: int):
        """DataFrame(word, similarity) (ML:375-393)."""
        res = self.findSynonymsArray(word_or_vec, num)
        return frames.make_frame({"word": [w for w, _ in res], "similarity": [s for _, s in res]})

    # ------------------------------------------------------------- transform
    def transform(self, dataset):
        """Append the average of the in-vocabulary word vectors of each sentence
        (OOV words dropped, empty -> zero vector) as the last column (ML:432-460)."""
        self._validate_input(dataset)
        h = self._require_handle()
        sents = frames.get_column(dataset, self.getInputCol())
        get = self._index.get
        d = self._vsize
        out: List[np.ndarray] = []
        step = 1 << 16
        for lo in range(0, len(sents), step):
            flat: List[int] = []
            offs = [0]
            for s in sents[lo:lo + step]:
                if s is not None:
                    flat.extend(i for i in (get(w) for w in s) if i is not None)
                offs.append(len(flat))
            avg = h.pull_average(np.asarray(flat, np.int64), np.asarray(offs, np.int64))
            out.extend(avg[i].astype(np.float64) for i in range(avg.shape[0]))
        if not sents:
            out = []
        return frames.append_column(dataset, self.getOutputCol(), out)

    # ------------------------------------------------------------- persistence
    def write(self) -> _Writer:
        return _Writer(self)

    def save(self, path: str):
        self.write().save(path)

    def _save_impl(self, path: str):
        h = self._require_handle()
        if self._is_writer_rank():
            os.makedirs(path, exist_ok=True)
            # the host actually used is what a later load() must reconnect to (ML:516)
            self._save_metadata(path, JAVA_MODEL_CLASS)
            wdir = os.path.join(path, "words")
            os.makedirs(wdir, exist_ok=True)
            with open(os.path.join(wdir, "part-00000"), "w", encoding="utf-8", newline="\n") as f:
                for i in range(self.numWords):
                    f.write(self._words[i])
                    f.write("\n")
            with open(os.path.join(wdir, "_SUCCESS"), "w"):
                pass
        h.save(os.path.abspath(path), {"formatVersion": FORMAT_VERSION})

    @classmethod
    def read(cls):
        return _Reader(cls)

    @classmethod
    def load(cls, path: str, parameterServerHost: Optional[str] = None,
             parameterServerConfig: Optional[dict] = None) -> "ServerSideGlintWord2VecModel":
        """``load(path)`` / ``load(path, host)`` / ``load(path, host, config)``
        (ML:573,584-586,597-599; PY:353-373).  A given host/config overrides the
        saved ``parameterServerHost``/``parameterServerConfig``."""
        return _Reader(cls).load(path, parameterServerHost, parameterServerConfig)

    # ------------------------------------------------------------- lifecycle
    def toLocal(self):
        """Pull everything into a stock local ``Word2VecModel`` (ML:483, MLLIB:651-654)."""
        from .local_model import Word2VecModel
        h = self._require_handle()
        mat = h.pull(np.arange(self.numWords, dtype=np.int64))
        local = Word2VecModel(words=[self._words[i] for i in range(self.numWords)], vectors=mat)
        for name in ("inputCol", "outputCol", "vectorSize", "windowSize", "numPartitions", "minCount",
                     "maxSentenceLength", "stepSize", "maxIter", "seed"):
            if self.isDefined(name) and local.hasParam(name):
                local.set(name, self.getOrDefault(name))
        return local

    def stop(self, terminateOtherClients: bool = False):
        """Destroy the distributed matrix and stop an integrated server group
        (ML:492-496, MLLIB:664-667; PY:375-383)."""
        if self._handle is not None:
            self._handle.destroy()
            self._handle.terminate(terminateOtherClients)
            self._handle = None

    def copy(self, extra=None):
        that = ServerSideGlintWord2VecModel(self.uid, self._words, self._handle, self._index)
        self._copyValues(that, extra)
        that._parent = self._parent
        that.trainingReport = self.trainingReport
        return that

    # the MLlib model is Serializable and usable inside closures (SPEC:230,250);
    # here: picklable when attached to a separate / spawned server group.
    def __getstate__(self):
        st = dict(self.__dict__)
        h = st.get("_handle")
        if h is not None and not isinstance(h, _cluster.RemoteHandle):
            raise TypeError("an in-process model cannot be pickled; use a server group (parameterServerHost)")
        if h is not None:
            st["_handle"] = ("remote", h.addr, h.matrix_id, h.host, h.cols)
        return st

    def __setstate__(self, st):
        h = st.get("_handle")
        self.__dict__.update(st)
        if isinstance(h, tuple):
            handle = _cluster.RemoteHandle(h[1][0], h[1][1], matrix_id=h[2], persist_host=h[3])
            handle.cols = h[4]
            self._handle = handle


class _Reader:
    def __init__(self, cls):
        self.cls = cls

    def load(self, path: str, host: Optional[str] = None, config: Optional[dict] = None):
        from .estimator import engine_options_from_params, open_handle_for_load
        meta = ServerSideGlintWord2VecBase._load_metadata(path, JAVA_MODEL_CLASS)
        with open(os.path.join(path, "words", "part-00000"), encoding="utf-8", newline="\n") as f:
            text = f.read()
        words = text.split("\n")
        if words and words[-1] == "":
            words.pop()                               # final newline; interior empty lines are words (Q9)
        model = self.cls(uid=meta["uid"], words=words)
        model._get_and_set_params(meta)
        # overrides (ML:542-544)
        if host is not None and host != "":
            model.set("parameterServerHost", host)
        if config:
            model.set("parameterServerConfig", config)
        opts = engine_options_from_params(model)
        model._handle = open_handle_for_load(os.path.abspath(path), model.getParameterServerHost(),
                                             model.getNumParameterServers(), opts)
        return model
