"""Stock (non-distributed) ``Word2VecModel`` -- the target of ``toLocal``.

Reference: ``toLocal`` converts the distributed model into Spark's default
``org.apache.spark.ml.feature.Word2VecModel`` (ML:483, MLLIB:651-654), which
the spec then saves (SPEC:400-415).  Saved layout = public Spark 2.4 format:
``<path>/metadata/part-00000`` + ``<path>/data/part-00000.parquet`` with schema
``word: string, vector: array<float>``, written with
pyarrow.
"""
from __future__ import annotations

import os
import shutil
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import frames
from .params import Params, _to_float, _to_int, _to_str

JAVA_LOCAL_CLASS = "org.apache.spark.ml.feature.Word2VecModel"


class Word2VecModel(Params):
    _uid_prefix = "w2v"

    def __init__(self, uid: Optional[str] = None, words: Optional[Sequence[str]] = None,
                 vectors: Optional[np.ndarray] = None):
        super().__init__(uid)
        d = self._declare
        d("inputCol", "input column name.", has_default=False, converter=_to_str)
        d("outputCol", "output column name.", has_default=False, converter=_to_str)
        d("seed", "random seed.", -1961189076, converter=_to_int)
        d("stepSize", "Step size (learning rate)", 0.025, None, _to_float)
        d("maxIter", "maximum number of iterations (>= 0).", 1, None, _to_int)
        d("vectorSize", "the dimension of codes after transforming from words (> 0)", 100, None, _to_int)
        d("windowSize", "the window size (> 0)", 5, None, _to_int)
        d("numPartitions", "number of partitions for sentences of words (> 0)", 1, None, _to_int)
        d("minCount", "minimum token count (>= 0)", 5, None, _to_int)
        d("maxSentenceLength", "Maximum length (in words) of each sentence (> 0)", 1000, None, _to_int)
        self._words = list(words) if words is not None else []
        self._vectors = np.asarray(vectors, dtype=np.float32) if vectors is not None \
            else np.zeros((0, 0), np.float32)
        self._index = {w: i for i, w in enumerate(self._words)}
        self._norms = None

    def setInputCol(self, v): return self.set("inputCol", v)
    def setOutputCol(self, v): return self.set("outputCol", v)
    def getInputCol(self): return self.getOrDefault("inputCol")
    def getOutputCol(self): return self.getOrDefault("outputCol")
    def getVectorSize(self): return int(self._vectors.shape[1])

    @property
    def numWords(self) -> int:
        return len(self._words)

    def getVectors(self):
        return frames.make_frame({"word": list(self._words),
                                  "vector": [v.astype(np.float64) for v in self._vectors]})

    def findSynonymsArray(self, word_or_vec: Union[str, Sequence[float]], num: int) -> List[Tuple[str, float]]:
        if num <= 0:
            raise ValueError("Number of similar words should > 0")
        if isinstance(word_or_vec, str):
            if word_or_vec not in self._index:
                raise KeyError(f"{word_or_vec} not in vocabulary")
            q = self._vectors[self._index[word_or_vec]]
            wopt = word_or_vec
        else:
            q = np.asarray(word_or_vec, np.float32)
            wopt = None
        if self._norms is None:
            self._norms = np.linalg.norm(self._vectors, axis=1)
        qn = np.linalg.norm(q)
        q = q / qn if qn > 0 else q
        with np.errstate(divide="ignore", invalid="ignore"):
            cos = np.where(self._norms > 0, (self._vectors @ q) / self._norms, 0.0)
        k = min(num + 1, self.numWords)
        top = np.argpartition(-cos, k - 1)[:k]
        top = top[np.argsort(-cos[top], kind="stable")]
        res = [(self._words[i], float(cos[i])) for i in top if self._words[i] != wopt]
        return res[:num]

    def findSynonyms(self, word_or_vec, num: int):
        res = self.findSynonymsArray(word_or_vec, num)
        return frames.make_frame({"word": [w for w, _ in res], "similarity": [s for _, s in res]})

    def transform(self, dataset):
        sents = frames.get_column(dataset, self.getInputCol())
        d = self.getVectorSize()
        out = []
        for s in sents:
            idx = [self._index[w] for w in (s or []) if w in self._index]
            out.append(self._vectors[idx].mean(0).astype(np.float64) if idx else np.zeros(d))
        return frames.append_column(dataset, self.getOutputCol(), out)

    # -- persistence
    def save(self, path: str):
        if os.path.exists(path):
            raise IOError(f"Path {path} already exists. To overwrite it, please use write.overwrite().save(path).")
        self._save_impl(path)

    def write(self):
        m = self

        class _W:
            _ow = False

            def overwrite(self):
                self._ow = True
                return self

            def save(self, path):
                if os.path.exists(path) and self._ow:
                    shutil.rmtree(path)
                m.save(path)
        return _W()

    def _save_impl(self, path: str):
        import pyarrow as pa
        import pyarrow.parquet as pq
        self._save_metadata(path, JAVA_LOCAL_CLASS)
        ddir = os.path.join(path, "data")
        os.makedirs(ddir, exist_ok=True)
        table = pa.table({"word": pa.array(self._words, type=pa.string()),
                          "vector": pa.array([v.tolist() for v in self._vectors], type=pa.list_(pa.float32()))})
        pq.write_table(table, os.path.join(ddir, "part-00000.parquet"))
        with open(os.path.join(ddir, "_SUCCESS"), "w"):
            pass

    @classmethod
    def load(cls, path: str) -> "Word2VecModel":
        import pyarrow.parquet as pq
        meta = cls._load_metadata(path, JAVA_LOCAL_CLASS)
        ddir = os.path.join(path, "data")
        files = sorted(f for f in os.listdir(ddir) if f.endswith(".parquet"))
        words: List[str] = []
        vecs = []
        for f in files:
            t = pq.read_table(os.path.join(ddir, f))
            words.extend(t.column("word").to_pylist())
            vecs.extend(t.column("vector").to_pylist())
        m = cls(uid=meta["uid"], words=words, vectors=np.asarray(vecs, np.float32))
        m._get_and_set_params(meta)
        return m
