"""MLlib-flavoured API (``org.apache.spark.mllib.feature.ServerSideGlintWord2Vec``).

The reference has two public layers: the ML estimator delegates everything to
this builder-style MLlib trainer (MLLIB:65-449) and model (MLLIB:460-726).
Here both layers share one implementation; this module only provides the MLlib
names (``setLearningRate``/``setNumIterations``, ``transform(word)`` raising on
OOV, ``getVectors`` as a map, ``save(path)``/``load(path[, host[, config]])``).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from .estimator import ServerSideGlintWord2Vec as _MLEstimator
from .model import ServerSideGlintWord2VecModel as _MLModel


class MLlibServerSideGlintWord2VecModel:
    """MLLIB:460-669."""

    formatVersion = "1.0"                          # MLLIB:488

    def __init__(self, ml_model: _MLModel):
        self._m = ml_model

    @property
    def numWords(self) -> int:                     # MLLIB:468
        return self._m.numWords

    @property
    def vectorSize(self) -> int:                   # MLLIB:473
        return self._m._vsize

    @property
    def wordList(self) -> List[str]:               # MLLIB:478-481: words ordered by row index
        return list(self._m._words)

    def transform(self, word_or_words: Union[str, Iterable[str]]):
        """``transform(word)`` -> vector; ``transform(iterator)`` -> iterator of
        vectors in 10 000-word batches; both raise on OOV (MLLIB:511-543)."""
        if isinstance(word_or_words, str):
            try:
                return self._m.transformWord(word_or_words)
            except KeyError:
                raise KeyError(f"{word_or_words} not in vocabulary") from None
        return self._m.transformWords(word_or_words)

    def findSynonyms(self, word_or_vec, num: int) -> List[Tuple[str, float]]:   # MLLIB:554-572
        return self._m.findSynonymsArray(word_or_vec, num)

    def getVectors(self) -> Dict[str, np.ndarray]:                               # MLLIB:638-641
        return self._m.getVectorsMap()

    def toLocal(self):                                                           # MLLIB:651-654
        return self._m.toLocal()

    def save(self, path: str):                                                   # MLLIB:493-498
        self._m.save(path)

    def stop(self, terminateOtherClients: bool = False):                         # MLLIB:664-667
        self._m.stop(terminateOtherClients)

    @classmethod
    def load(cls, path: str, parameterServerHost: str = "", parameterServerConfig: Optional[dict] = None):
        """MLLIB:683,696,710."""
        return cls(_MLModel.load(path, parameterServerHost, parameterServerConfig))

    @property
    def ml(self) -> _MLModel:
        return self._m


class MLlibServerSideGlintWord2Vec:
    """Builder-style trainer with the 15 MLlib setters (MLLIB:92-244)."""

    def __init__(self):
        self._est = _MLEstimator()
        # MLlib defaults that differ from the ML layer: learningRate 0.01875 is the same; seed random (MLLIB:71)
        self._est.setSeed(int(np.random.SeedSequence().entropy % (2 ** 31)))

    def _chk(self, cond, msg):
        if not cond:
            raise ValueError("requirement failed: " + msg)

    def setMaxSentenceLength(self, v):
        self._chk(v > 0, f"Maximum length of sentences must be positive but got {v}")
        self._est.setMaxSentenceLength(v); return self

    def setVectorSize(self, v):
        self._chk(v > 0, f"vector size must be positive but got {v}")
        self._est.setVectorSize(v); return self

    def setLearningRate(self, v):
        self._chk(v > 0, f"Initial learning rate must be positive but got {v}")
        self._est.setStepSize(v); return self

    def setNumPartitions(self, v):
        self._chk(v > 0, f"Number of partitions must be positive but got {v}")
        self._est.setNumPartitions(v); return self

    def setNumIterations(self, v):
        self._chk(v >= 0, f"Number of iterations must be nonnegative but got {v}")
        self._est.setMaxIter(v); return self

    def setSeed(self, v):
        self._est.setSeed(v); return self

    def setWindowSize(self, v):
        self._chk(v > 0, f"Window of words must be positive but got {v}")
        self._est.setWindowSize(v); return self

    def setMinCount(self, v):
        self._chk(v >= 0, f"Minimum number of times must be nonnegative but got {v}")
        self._est.setMinCount(v); return self

    def setBatchSize(self, v):
        self._chk(v > 0, f"Mini batch size must be positive but got {v}")
        self._est.setBatchSize(v); return self

    def setN(self, v):
        self._chk(v > 0, f"Number of negative examples must be positive but got {v}")
        self._est.setN(v); return self

    def setSubsampleRatio(self, v):
        self._chk(v >= 0, f"Subsample ratio must be nonnegative but got {v}")
        self._est.setSubsampleRatio(v); return self

    def setNumParameterServers(self, v):
        self._chk(v > 0, f"Number of parameter servers must be positive but got {v}")
        self._est.setNumParameterServers(v); return self

    def setParameterServerHost(self, v):
        self._est.setParameterServerHost(v); return self

    def setParameterServerConfig(self, v):
        self._est.setParameterServerConfig(v); return self

    def setUnigramTableSize(self, v):
        self._chk(v > 0, f"Unigram table size must be positive but got {v}")
        self._est.setUnigramTableSize(v); return self

    def fit(self, sentences: Iterable[Sequence[str]]) -> MLlibServerSideGlintWord2VecModel:
        """``fit(RDD[Iterable[String]])`` (MLLIB:310-326)."""
        return MLlibServerSideGlintWord2VecModel(self._est.fit(sentences))


# the MLlib layer's own class names (org.apache.spark.mllib.feature.*), MLLIB:65,460
ServerSideGlintWord2Vec = MLlibServerSideGlintWord2Vec
ServerSideGlintWord2VecModel = MLlibServerSideGlintWord2VecModel
