"""A small re-implementation of Spark ML's ``Param``/``Params`` machinery and
the parameter trait of the reference (component C13, ML:40-222, mirrored in
Python at PY:101-170).

Names, defaults, docs and validators follow ``ServerSideGlintWord2VecBase``:

=========================  ==========  ======================================
param                      default     cite
=========================  ==========  ======================================
vectorSize                 100         ML:48-54
windowSize                 5           ML:61-67
numPartitions              1           ML:74-80
minCount                   5           ML:88-93
maxSentenceLength          1000        ML:102-108
batchSize                  50          ML:116-120
n                          5           ML:128-132
subsampleRatio             1e-6        ML:141-146
numParameterServers        5           ML:154-159
parameterServerHost        ""          ML:168-174
parameterServerConfig      {}          ML:183-195 (JSON codec)
unigramTableSize           100000000   ML:204-209
stepSize                   0.01875     ML:211
maxIter                    1           ML:212
=========================  ==========  ======================================
"""
from __future__ import annotations

import copy as _copy
import json
import os
import time
import uuid
from typing import Any, Callable, Dict, Optional


class Param:
    def __init__(self, parent: str, name: str, doc: str, validator: Optional[Callable[[Any], bool]] = None,
                 converter: Optional[Callable[[Any], Any]] = None):
        self.parent = parent
        self.name = name
        self.doc = doc
        self.validator = validator
        self.converter = converter

    def __repr__(self):
        return f"Param(parent={self.parent!r}, name={self.name!r})"

    def __hash__(self):
        return hash(self.name)

    def __eq__(self, other):
        return isinstance(other, Param) and other.name == self.name


def _gt0(v):
    return v > 0


def _ge0(v):
    return v >= 0


def _to_int(v):
    if isinstance(v, bool) or (isinstance(v, float) and int(v) != v):
        raise TypeError(f"Could not convert {v!r} to int")
    return int(v)


def _to_float(v):
    return float(v)


def _to_str(v):
    if not isinstance(v, str):
        raise TypeError(f"Could not convert {v!r} to string")
    return v


def _to_config(v):
    """``parameterServerConfig``: a (possibly nested) dict; JSON text and dotted
    keys are accepted and normalised to a flat dotted-key dict, which is what
    the reference's ``toJavaPathMap`` produces on load (ML:548-560)."""
    if v is None:
        return {}
    if isinstance(v, str):
        v = json.loads(v) if v.strip() else {}
    if not isinstance(v, dict):
        raise TypeError("parameterServerConfig must be a dict or JSON object string")
    return flatten_config(v)


def flatten_config(d: dict, prefix: str = "") -> dict:
    out = {}
    for k, v in d.items():
        key = f"{prefix}.{k}" if prefix else str(k)
        if isinstance(v, dict):
            out.update(flatten_config(v, key))
        else:
            out[key] = v
    return out


def nest_config(flat: dict) -> dict:
    """Inverse of ``flatten_config`` -- the nested JSON object the reference
    writes into the metadata (ML:187-189)."""
    out: dict = {}
    for k, v in flat.items():
        parts = k.split(".")
        cur = out
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = v
    return out


class Params:
    """Spark-ML style parameter container."""

    _uid_prefix = "params"

    def __init__(self, uid: Optional[str] = None):
        self.uid = uid or f"{self._uid_prefix}_{uuid.uuid4().hex[:12]}"
        self._paramMap: Dict[Param, Any] = {}
        self._defaultParamMap: Dict[Param, Any] = {}
        self._params: Dict[str, Param] = {}

    # -- declaration
    def _declare(self, name, doc, default=None, validator=None, converter=None, has_default=True) -> Param:
        p = Param(self.uid, name, doc, validator, converter)
        self._params[name] = p
        setattr(self, name, p)
        if has_default:
            self._defaultParamMap[p] = default
        return p

    # -- Spark API
    @property
    def params(self):
        return [self._params[k] for k in sorted(self._params)]

    def hasParam(self, name: str) -> bool:
        return name in self._params

    def getParam(self, name: str) -> Param:
        if name not in self._params:
            raise ValueError(f"Param {name} does not exist.")
        return self._params[name]

    def _resolve(self, param) -> Param:
        return self.getParam(param) if isinstance(param, str) else self.getParam(param.name)

    def isSet(self, param) -> bool:
        return self._resolve(param) in self._paramMap

    def hasDefault(self, param) -> bool:
        return self._resolve(param) in self._defaultParamMap

    def isDefined(self, param) -> bool:
        return self.isSet(param) or self.hasDefault(param)

    def getOrDefault(self, param):
        p = self._resolve(param)
        if p in self._paramMap:
            return self._paramMap[p]
        if p in self._defaultParamMap:
            return self._defaultParamMap[p]
        raise KeyError(f"Failed to find a default value for {p.name}")

    def getDefault(self, param):
        return self._defaultParamMap.get(self._resolve(param))

    def set(self, param, value):
        p = self._resolve(param)
        if p.converter is not None:
            value = p.converter(value)
        if p.validator is not None and not p.validator(value):
            raise ValueError(f"{self.uid} parameter {p.name} given invalid value {value!r}.")
        self._paramMap[p] = value
        return self

    _set_one = set

    def _set(self, **kwargs):
        for k, v in kwargs.items():
            if v is not None:
                self.set(k, v)
        return self

    def _setDefault(self, **kwargs):
        for k, v in kwargs.items():
            self._defaultParamMap[self.getParam(k)] = v
        return self

    def clear(self, param):
        self._paramMap.pop(self._resolve(param), None)
        return self

    def explainParam(self, param) -> str:
        p = self._resolve(param)
        bits = []
        if self.hasDefault(p):
            bits.append(f"default: {self._defaultParamMap[p]}")
        if self.isSet(p):
            bits.append(f"current: {self._paramMap[p]}")
        if not bits:
            bits.append("undefined")
        return f"{p.name}: {p.doc} ({', '.join(bits)})"

    def explainParams(self) -> str:
        return "\n".join(self.explainParam(p) for p in self.params)

    def extractParamMap(self, extra: Optional[dict] = None) -> dict:
        m = dict(self._defaultParamMap)
        m.update(self._paramMap)
        if extra:
            for k, v in extra.items():
                m[self._resolve(k)] = v
        return m

    def copy(self, extra: Optional[dict] = None):
        that = _copy.copy(self)
        that._paramMap = dict(self._paramMap)
        that._defaultParamMap = dict(self._defaultParamMap)
        if extra:
            for k, v in extra.items():
                that.set(k, v)
        return that

    def _copyValues(self, to: "Params", extra: Optional[dict] = None):
        """``copyValues``: explicitly set + default values for params ``to`` also has."""
        for p, v in self._defaultParamMap.items():
            if to.hasParam(p.name) and not to.hasDefault(p.name):
                to._defaultParamMap[to.getParam(p.name)] = v
        for p, v in self._paramMap.items():
            if to.hasParam(p.name):
                to._paramMap[to.getParam(p.name)] = v
        if extra:
            for k, v in extra.items():
                to.set(k, v)
        return to

    # -- metadata (DefaultParamsWriter / DefaultParamsReader, ML:505, ML:514)
    def _json_value(self, p: Param, v):
        if p.name == "parameterServerConfig":
            return nest_config(v)
        return v

    def _metadata(self, cls_name: str) -> dict:
        return {
            "class": cls_name,
            "timestamp": int(time.time() * 1000),
            "sparkVersion": "2.4.0",
            "uid": self.uid,
            "paramMap": {p.name: self._json_value(p, v) for p, v in self._paramMap.items()},
            "defaultParamMap": {p.name: self._json_value(p, v) for p, v in self._defaultParamMap.items()},
        }

    def _save_metadata(self, path: str, cls_name: str, extra: Optional[dict] = None):
        mdir = os.path.join(path, "metadata")
        os.makedirs(mdir, exist_ok=True)
        meta = self._metadata(cls_name)
        if extra:
            meta.update(extra)
        with open(os.path.join(mdir, "part-00000"), "w", encoding="utf-8") as f:
            f.write(json.dumps(meta, separators=(",", ":"), ensure_ascii=False) + "\n")
        with open(os.path.join(mdir, "_SUCCESS"), "w"):
            pass

    @staticmethod
    def _load_metadata(path: str, expected_class: Optional[str] = None) -> dict:
        with open(os.path.join(path, "metadata", "part-00000"), encoding="utf-8") as f:
            meta = json.loads(f.readline())
        if expected_class is not None and meta.get("class") != expected_class:
            raise ValueError(f"Error loading metadata: Expected class name {expected_class} "
                             f"but found class name {meta.get('class')}")
        return meta

    def _get_and_set_params(self, meta: dict, skip=()):
        """``metadata.getAndSetParams`` (ML:542)."""
        for name, v in meta.get("defaultParamMap", {}).items():
            if self.hasParam(name) and name not in skip:
                p = self.getParam(name)
                self._defaultParamMap[p] = p.converter(v) if p.converter else v
        for name, v in meta.get("paramMap", {}).items():
            if self.hasParam(name) and name not in skip:
                self.set(name, v)
        return self


# ----------------------------------------------------------------------------

def java_string_hash(s: str) -> int:
    """``String.hashCode`` of the JVM (signed 32-bit)."""
    h = 0
    for c in s:
        h = (31 * h + ord(c)) & 0xFFFFFFFF
    return h - (1 << 32) if h & 0x80000000 else h


class ServerSideGlintWord2VecBase(Params):
    """Params shared by the estimator and the model (``ServerSideGlintWord2VecBase``, ML:40-222)."""

    _java_class = "org.apache.spark.ml.feature.ServerSideGlintWord2Vec"

    def _declare_w2v_params(self):
        d = self._declare
        d("inputCol", "input column name.", has_default=False, converter=_to_str)
        d("outputCol", "output column name.", has_default=False, converter=_to_str)
        # HasSeed default: this.getClass.getName.hashCode (Java String.hashCode)
        d("seed", "random seed.", default=java_string_hash(self._java_class), converter=_to_int)
        d("stepSize", "Step size to be used for each iteration of optimization (> 0).", 0.01875,
          _gt0, _to_float)
        d("maxIter", "maximum number of iterations (>= 0).", 1, _ge0, _to_int)
        d("vectorSize", "the dimension of codes after transforming from words (> 0)", 100,
          _gt0, _to_int)
        d("windowSize", "the window size (context words from [-window, window]) (> 0)", 5,
          _gt0, _to_int)
        d("numPartitions", "number of partitions for sentences of words (> 0)", 1, _gt0, _to_int)
        d("minCount", "the minimum number of times a token must appear to be included in the "
          "word2vec model's vocabulary (>= 0)", 5, _ge0, _to_int)
        d("maxSentenceLength", "Maximum length (in words) of each sentence in the input data. Any sentence "
          "longer than this threshold will be divided into chunks up to the size (> 0)", 1000,
          _gt0, _to_int)
        d("batchSize", "the mini batch size", 50, None, _to_int)
        d("n", "the number of random negative examples", 5, None, _to_int)
        d("subsampleRatio", "the ratio controlling how much subsampling occurs. "
          "Smaller values mean frequent words are less likely to be kept", 1e-6, None, _to_float)
        d("numParameterServers", "the number of parameter servers to create (here: column shards = GPUs)",
          5, None, _to_int)
        d("parameterServerHost", "the master host of the running parameter servers. If this is not set a "
          "standalone parameter server cluster is started in this application.", "", None, _to_str)
        d("parameterServerConfig", "the parameter server configuration (engine options).", {}, None, _to_config)
        d("unigramTableSize", "the size of the unigram table. Only needs to be changed to a lower value "
          "if there is not enough memory for local testing", 100000000, None, _to_int)

    # getters (ML:54,67,80,93,108,120,132,146,159,174,209)
    def getInputCol(self): return self.getOrDefault("inputCol")
    def getOutputCol(self): return self.getOrDefault("outputCol")
    def getSeed(self): return self.getOrDefault("seed")
    def getStepSize(self): return self.getOrDefault("stepSize")
    def getMaxIter(self): return self.getOrDefault("maxIter")
    def getVectorSize(self): return self.getOrDefault("vectorSize")
    def getWindowSize(self): return self.getOrDefault("windowSize")
    def getNumPartitions(self): return self.getOrDefault("numPartitions")
    def getMinCount(self): return self.getOrDefault("minCount")
    def getMaxSentenceLength(self): return self.getOrDefault("maxSentenceLength")
    def getBatchSize(self): return self.getOrDefault("batchSize")
    def getN(self): return self.getOrDefault("n")
    def getSubsampleRatio(self): return self.getOrDefault("subsampleRatio")
    def getNumParameterServers(self): return self.getOrDefault("numParameterServers")
    def getParameterServerHost(self): return self.getOrDefault("parameterServerHost")
    def getParameterServerConfig(self): return dict(self.getOrDefault("parameterServerConfig"))
    def getUnigramTableSize(self): return self.getOrDefault("unigramTableSize")

    # -- schema handling (validateAndTransformSchema, ML:217-221)
    def _validate_input(self, data):
        """Input column must hold arrays of strings; returns the column names."""
        from .frames import column_names, first_non_null
        cols = column_names(data)
        if cols is None:
            return None
        ic = self.getInputCol()
        if ic not in cols:
            raise ValueError(f"Field \"{ic}\" does not exist. Available fields: {', '.join(cols)}")
        sample = first_non_null(data, ic)
        if sample is not None:
            if isinstance(sample, (str, bytes)) or not hasattr(sample, "__iter__"):
                raise TypeError(f"Column {ic} must be of type array<string> but was actually {type(sample).__name__}.")
            for w in sample:
                if not isinstance(w, str):
                    raise TypeError(f"Column {ic} must be of type array<string> but holds {type(w).__name__} elements.")
                break
        oc = self.getOutputCol() if self.isDefined("outputCol") else None
        if oc is not None and oc in cols:
            raise ValueError(f"Column {oc} already exists.")
        return cols

    def transformSchema(self, schema):
        """Schema = ordered list of (name, type) pairs; appends the vector column."""
        names = [n for n, _ in schema]
        ic = self.getInputCol()
        if ic not in names:
            raise ValueError(f"Field \"{ic}\" does not exist.")
        typ = dict(schema)[ic]
        if typ not in ("array<string>", "ArrayType(StringType,true)", "ArrayType(StringType,false)"):
            raise TypeError(f"Column {ic} must be of type array<string> but was actually {typ}.")
        oc = self.getOutputCol()
        if oc in names:
            raise ValueError(f"Column {oc} already exists.")
        return list(schema) + [(oc, "vector")]
