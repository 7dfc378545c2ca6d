"""API-surface parity with the reference's ML / MLlib / Python layers
(SURVEY.md Appendix A): params, defaults, validators, persistence, model ops."""
import json
import os
import pickle

import numpy as np
import pandas as pd
import pytest

from glint_word2vec_b200 import (MLlibServerSideGlintWord2Vec, MLlibServerSideGlintWord2VecModel,
                                 ServerSideGlintWord2Vec, ServerSideGlintWord2VecModel, Word2VecModel)
from glint_word2vec_b200.api.params import flatten_config, java_string_hash, nest_config

CPU = {"device": "cpu"}


def test_param_defaults_match_reference():
    est = ServerSideGlintWord2Vec()
    expect = dict(vectorSize=100, windowSize=5, numPartitions=1, minCount=5, maxSentenceLength=1000,
                  batchSize=50, n=5, subsampleRatio=1e-6, numParameterServers=5, parameterServerHost="",
                  unigramTableSize=100000000, stepSize=0.01875, maxIter=1)           # ML:48-212
    for k, v in expect.items():
        assert est.getOrDefault(k) == v, k
        getter = "get" + k[0].upper() + k[1:]
        assert getattr(est, getter)() == v
    assert est.getParameterServerConfig() == {}
    assert est.uid.startswith("gw2v")                                               # ML:231
    assert est.getSeed() == java_string_hash("org.apache.spark.ml.feature.ServerSideGlintWord2Vec")
    assert not est.isDefined("inputCol")
    assert "vectorSize" in est.explainParams() and "default: 100" in est.explainParam("vectorSize")


def test_setters_validators_and_copy():
    est = (ServerSideGlintWord2Vec().setVectorSize(32).setWindowSize(3).setStepSize(0.05).setNumPartitions(2)
           .setMaxIter(2).setSeed(7).setMinCount(1).setMaxSentenceLength(10).setBatchSize(8).setN(3)
           .setSubsampleRatio(1e-3).setNumParameterServers(2).setParameterServerHost("")
           .setParameterServerConfig({"a": {"b": 1}}).setUnigramTableSize(1000)
           .setInputCol("s").setOutputCol("o"))
    assert est.getVectorSize() == 32 and est.getN() == 3 and est.getParameterServerConfig() == {"a.b": 1}
    for name, bad in (("vectorSize", 0), ("windowSize", 0), ("numPartitions", 0), ("minCount", -1),
                      ("maxSentenceLength", 0), ("stepSize", 0.0), ("maxIter", -1)):
        with pytest.raises(ValueError):
            est.set(name, bad)
    with pytest.raises(TypeError):
        est.setVectorSize(1.5)
    c = est.copy({"vectorSize": 64})
    assert c.getVectorSize() == 64 and est.getVectorSize() == 32 and c.getN() == 3 and c.uid == est.uid
    with pytest.raises(TypeError):
        ServerSideGlintWord2Vec(noSuchParam=1)
    # fit-time validation of the params the reference leaves unvalidated at the ML level (Q7)
    with pytest.raises(ValueError):
        ServerSideGlintWord2Vec(batchSize=0, inputCol="s", outputCol="o").fit({"s": [["a"] * 10]})


def test_config_codec_roundtrip():
    nested = {"glint": {"master": {"port": 13380}}, "akka": {"remote": {"artery": {"canonical": {"port": 13381}}}}}
    flat = flatten_config(nested)
    assert flat == {"glint.master.port": 13380, "akka.remote.artery.canonical.port": 13381}   # cf. SEPCONF:3-11
    assert nest_config(flat) == nested


def test_estimator_persistence(tmp_path):
    est = ServerSideGlintWord2Vec(vectorSize=17, seed=3, inputCol="a", outputCol="b",
                                  parameterServerConfig={"subsample_mode": "reference"})
    p = str(tmp_path / "est")
    est.save(p)
    meta = json.loads(open(os.path.join(p, "metadata", "part-00000")).read())
    assert meta["class"] == "org.apache.spark.ml.feature.ServerSideGlintWord2Vec"
    assert meta["paramMap"]["vectorSize"] == 17 and meta["defaultParamMap"]["vectorSize"] == 100
    assert os.path.exists(os.path.join(p, "metadata", "_SUCCESS"))
    est2 = ServerSideGlintWord2Vec.load(p)
    assert est2.uid == est.uid and est2.getVectorSize() == 17 and est2.getInputCol() == "a"
    assert est2.getParameterServerConfig() == {"subsample_mode": "reference"}
    with pytest.raises(IOError):
        est.save(p)
    est.write().overwrite().save(p)


@pytest.fixture(scope="module")
def small_model():
    from glint_word2vec_b200.data.synthetic import synthetic_capitals_corpus
    sents = synthetic_capitals_corpus(3000, seed=1)
    est = ServerSideGlintWord2Vec(vectorSize=24, seed=5, stepSize=0.05, maxIter=3, minCount=2,
                                  numParameterServers=1, inputCol="sentence", outputCol="vec",
                                  parameterServerConfig=dict(CPU, subsample_mode="reference"))
    model = est.fit(pd.DataFrame({"sentence": sents}))
    yield model
    model.stop()


def test_model_basics(small_model):
    m = small_model
    assert m.numWords == len(set(w for w in m._words)) > 300
    assert m.getVectorSize() == 24 and m.parent is not None and m.getOutputCol() == "vec"
    assert m.trainingReport["pairs"] > 0
    gv = m.getVectors()
    assert list(gv.columns) == ["word", "vector"] and len(gv) == m.numWords             # SPEC:384-398
    assert len(gv["vector"][0]) == 24


def test_transform_appends_last_and_averages(small_model):
    m = small_model
    df = pd.DataFrame({"id": [1, 2, 3, 4], "sentence": [["wien"], ["wien", "berlin", "zzz-oov"], [], ["zzz-oov"]],
                       "other": ["a", "b", "c", "d"]})
    out = m.transform(df)
    assert list(out.columns) == ["id", "sentence", "other", "vec"]                        # SPEC:260-288
    vw, vb = m.transformWord("wien"), m.transformWord("berlin")
    assert np.allclose(out["vec"][0], vw) and np.allclose(out["vec"][1], (vw + vb) / 2, atol=1e-6)
    assert not out["vec"][2].any() and not out["vec"][3].any()                            # empty / all-OOV -> zeros
    assert out["vec"][0].dtype == np.float64 and len(out["vec"][0]) == 24
    with pytest.raises(ValueError):
        m.transform(pd.DataFrame({"nope": [["a"]]}))
    with pytest.raises(TypeError):
        m.transform(pd.DataFrame({"sentence": ["not a list"]}))
    d = m.transform({"sentence": [["wien"]]})
    assert list(d.keys()) == ["sentence", "vec"]
    import pyarrow as pa
    t = m.transform(pa.table({"sentence": [["wien"], []]}))
    assert t.column_names == ["sentence", "vec"] and np.allclose(t.column("vec")[0].as_py(), vw)


def test_transform_schema(small_model):
    m = small_model
    assert m.transformSchema([("sentence", "array<string>"), ("x", "int")])[-1] == ("vec", "vector")
    with pytest.raises(TypeError):
        m.transformSchema([("sentence", "string")])
    with pytest.raises(ValueError):
        m.transformSchema([("sentence", "array<string>"), ("vec", "vector")])


def test_synonyms_semantics(small_model):
    m = small_model
    syn = m.findSynonymsArray("österreich", 5)
    assert len(syn) == 5 and all(w != "österreich" for w, _ in syn)                        # query word excluded
    assert all(syn[i][1] >= syn[i + 1][1] for i in range(4))
    v = m.transformWord("österreich")
    synv = m.findSynonymsArray(v, 5)
    assert synv[0][0] == "österreich" and abs(synv[0][1] - 1.0) < 1e-5                     # vector query keeps it
    df = m.findSynonyms("österreich", 3)
    assert list(df.columns) == ["word", "similarity"] and len(df) == 3                     # SPEC:307-325
    with pytest.raises(ValueError):
        m.findSynonymsArray("österreich", 0)                                               # MLLIB:587
    with pytest.raises(KeyError):
        m.findSynonymsArray("zzz-oov", 3)
    batch = m.findSynonymsArrayBatch(["österreich", v], 5)
    assert [w for w, _ in batch[0]] == [w for w, _ in syn] and batch[1][0][0] == "österreich"
    assert np.allclose([x for _, x in batch[0]], [x for _, x in syn], atol=1e-5)
    # cosine definition (MLLIB:589-617)
    mat = np.stack(m.getVectors()["vector"].to_list())
    cos = mat @ (v / np.linalg.norm(v)) / np.linalg.norm(mat, axis=1)
    assert abs(cos[m.wordIndex(syn[0][0])] - syn[0][1]) < 1e-5


def test_mllib_api(small_model):
    ml = MLlibServerSideGlintWord2VecModel(small_model)
    assert ml.numWords == small_model.numWords and ml.vectorSize == 24
    v = ml.transform("wien")
    assert v.shape == (24,) and v.any()
    with pytest.raises(KeyError, match="not in vocabulary"):
        ml.transform("zzz-oov")                                                            # MLLIB:516-517
    vs = list(ml.transform(iter(["wien", "berlin"])))
    assert len(vs) == 2 and np.allclose(vs[0], v)
    gv = ml.getVectors()
    assert isinstance(gv, dict) and len(gv) == ml.numWords and gv["wien"].dtype == np.float32
    assert ml.findSynonyms("wien", 2)[0][0] != "wien"
    b = MLlibServerSideGlintWord2Vec()
    for bad in (lambda: b.setVectorSize(0), lambda: b.setLearningRate(0), lambda: b.setNumIterations(-1),
                lambda: b.setWindowSize(0), lambda: b.setMinCount(-1), lambda: b.setBatchSize(0),
                lambda: b.setN(0), lambda: b.setNumParameterServers(0), lambda: b.setUnigramTableSize(0),
                lambda: b.setMaxSentenceLength(0), lambda: b.setNumPartitions(0), lambda: b.setSubsampleRatio(-1)):
        with pytest.raises(ValueError):
            bad()
    assert b.setNumIterations(0) is b                                                      # Q8: 0 iterations legal


def test_save_load_roundtrip_and_format(small_model, tmp_path):
    m = small_model
    p = str(tmp_path / "model")
    m.save(p)
    assert sorted(os.listdir(p)) == ["matrix", "metadata", "words"]
    meta = json.loads(open(os.path.join(p, "metadata", "part-00000")).readline())
    assert meta["class"] == "org.apache.spark.ml.feature.ServerSideGlintWord2VecModel"
    assert set(meta) >= {"class", "timestamp", "sparkVersion", "uid", "paramMap", "defaultParamMap"}
    assert meta["paramMap"]["parameterServerConfig"] == {"device": "cpu", "subsample_mode": "reference"}
    words = open(os.path.join(p, "words", "part-00000"), encoding="utf-8").read().split("\n")[:-1]
    assert words == list(m._words) and os.path.exists(os.path.join(p, "words", "_SUCCESS"))
    with pytest.raises(IOError):
        m.save(p)
    m2 = ServerSideGlintWord2VecModel.load(p)
    try:
        assert m2.uid == m.uid and m2.numWords == m.numWords
        for name in ("seed", "vectorSize", "stepSize", "maxIter", "minCount", "inputCol", "outputCol"):
            assert m2.getOrDefault(name) == m.getOrDefault(name)
        assert np.allclose(m2.transformWord("wien"), m.transformWord("wien"))
        assert m2.findSynonymsArray("wien", 3) == m.findSynonymsArray("wien", 3)
        m2.write().overwrite().save(p)
    finally:
        m2.stop()
    with pytest.raises(RuntimeError):
        m2.transformWord("wien")                                                           # stopped


def test_empty_word_survives_words_file(tmp_path):
    sents = [["", "a", "b", ""], ["a", "", "b"], ["", "a"]] * 5
    est = ServerSideGlintWord2Vec(vectorSize=8, minCount=1, seed=1, numParameterServers=1,
                                  inputCol="s", outputCol="o", parameterServerConfig=CPU)
    m = est.fit({"s": sents})
    p = str(tmp_path / "m")
    m.save(p)
    m2 = ServerSideGlintWord2VecModel.load(p)
    assert m2._words == list(m._words) and "" in m2._index                                 # Q9
    assert np.allclose(m2.transformWord(""), m.transformWord(""))
    m.stop(); m2.stop()


def test_to_local(small_model, tmp_path):
    loc = small_model.toLocal()
    assert isinstance(loc, Word2VecModel) and loc.uid.startswith("w2v") and loc.numWords == small_model.numWords
    a = small_model.findSynonymsArray("wien", 4)
    b = loc.findSynonymsArray("wien", 4)
    assert [w for w, _ in a] == [w for w, _ in b]
    p = str(tmp_path / "local")
    loc.save(p)                                                                             # SPEC:400-415
    import pyarrow.parquet as pq
    t = pq.read_table(os.path.join(p, "data", "part-00000.parquet"))
    assert t.column_names == ["word", "vector"] and str(t.schema.field("vector").type) in ("list<item: float>", "list<element: float>")
    meta = json.loads(open(os.path.join(p, "metadata", "part-00000")).readline())
    assert meta["class"] == "org.apache.spark.ml.feature.Word2VecModel"
    loc2 = Word2VecModel.load(p)
    assert loc2.findSynonymsArray("wien", 4)[0][0] == b[0][0]
    out = loc2.setInputCol("s").setOutputCol("v").transform(pd.DataFrame({"s": [["wien", "berlin"]]}))
    assert len(out["v"][0]) == 24


def test_in_process_model_refuses_pickle(small_model):
    with pytest.raises(TypeError):
        pickle.dumps(small_model)


def test_fit_accepts_plain_iterables_and_encoded():
    sents = [["a", "b", "c", "a", "b"]] * 30
    est = ServerSideGlintWord2Vec(vectorSize=8, minCount=1, seed=2, numParameterServers=1, parameterServerConfig=CPU)
    m = est.fit(iter(sents))
    assert m.numWords == 3
    m.stop()
    toks = np.array([0, 1, 2, 0, 1] * 30, dtype=np.int32)
    m2 = est.fitEncoded(toks, np.arange(0, 151, 5), np.array([60, 60, 30]))
    assert m2.numWords == 3 and m2._words[2] == "w2" and m2.transformWord("w1").shape == (8,)
    m2.stop()


def test_mllib_names_and_word_list(small_model):
    """The MLlib layer's own class names and `wordList` (MLLIB:65,460,478-481)."""
    from glint_word2vec_b200.api import mllib
    assert mllib.ServerSideGlintWord2Vec is mllib.MLlibServerSideGlintWord2Vec
    m = mllib.ServerSideGlintWord2VecModel(small_model)
    assert m.wordList == [w for w, _ in sorted(small_model._index.items(), key=lambda kv: kv[1])]
    assert m.formatVersion == "1.0" and m.vectorSize == small_model.getVectorSize()


def test_fit_text_file_equals_fit_on_sentences(tmp_path):
    """`fitTextFile` (native mmap loader) trains the same model as `fit` on the tokenised sentences."""
    from glint_word2vec_b200.data.synthetic import synthetic_capitals_corpus
    sentences = synthetic_capitals_corpus(seed=3)[:1500]
    path = tmp_path / "corpus.txt"
    path.write_text("\n".join(" ".join(s) for s in sentences), encoding="utf-8")
    kw = dict(inputCol="sentence", outputCol="vec", vectorSize=16, minCount=2, seed=5, numParameterServers=1,
              parameterServerConfig={"device": "cpu"})
    m1 = ServerSideGlintWord2Vec(**kw).fitTextFile(str(path))
    m2 = ServerSideGlintWord2Vec(**kw).fit(sentences)
    # streaming mode: the encoded corpus goes to disk block by block and is trained from a memory map
    kw3 = dict(kw, parameterServerConfig={"device": "cpu", "corpus_cache_dir": str(tmp_path / "cache")})
    m3 = ServerSideGlintWord2Vec(**kw3).fitTextFile(str(path))
    kw4 = dict(kw, parameterServerConfig={"device": "cpu", "stream_threshold_bytes": 0, "scratch_dir": str(tmp_path)})
    m4 = ServerSideGlintWord2Vec(**kw4).fitTextFile(str(path))
    try:
        assert m1.numWords == m2.numWords
        v1, v2, v3, v4 = m1.getVectorsMap(), m2.getVectorsMap(), m3.getVectorsMap(), m4.getVectorsMap()
        assert v1.keys() == v2.keys() == v3.keys()
        assert all(np.array_equal(v1[w], v2[w]) for w in list(v1)[:50])
        assert all(np.array_equal(v1[w], v3[w]) and np.array_equal(v1[w], v4[w]) for w in list(v1)[:50])
        assert (tmp_path / "cache" / "corpus.tokens.i32").exists()                 # explicit cache dir is kept
        assert not [p for p in tmp_path.iterdir() if p.name.startswith("gw2v-corpus-")]   # temporary one is removed
    finally:
        for m in (m1, m2, m3, m4):
            m.stop()


def test_tile_shared_negatives_keeps_the_quality_gates():
    """neg_sharing="tile" (library-GEMM path) learns the planted structure like the default mode."""
    from glint_word2vec_b200.data.synthetic import synthetic_capitals_corpus
    m = ServerSideGlintWord2Vec(inputCol="s", outputCol="v", vectorSize=100, stepSize=0.025, seed=1, minCount=5,
                                numParameterServers=1,
                                parameterServerConfig={"device": "cpu", "neg_sharing": "tile", "tile_centres": 64,
                                                       "tile_negatives": 32}).fit(synthetic_capitals_corpus())
    try:
        a, b = np.asarray(m.transformWord("österreich")), np.asarray(m.transformWord("wien"))
        assert float(a @ b / np.linalg.norm(a) / np.linalg.norm(b)) > 0.9
        q = np.asarray(m.transformWord("wien")) - a + np.asarray(m.transformWord("deutschland"))
        assert "berlin" in [w for w, _ in m.findSynonymsArray(q, 10)]
        assert m.getParameterServerConfig()["neg_sharing"] == "tile"
    finally:
        m.stop()


def test_every_module_compiles():
    """GPU-only modules are never imported by the CPU tier; a syntax error there must still fail here."""
    import glob
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "glint_word2vec_b200", "**", "*.py"), recursive=True)
    files += [os.path.join(root, f) for f in ("bench.py", "__graft_entry__.py", "benchmarks/bench_nn.py",
                                              "baseline/nccl_sgns.py", "scripts/run_integration.py")]
    assert len(files) > 20
    for f in files:
        py_compile.compile(f, doraise=True)


def test_command_line_train_query_export(tmp_path, capsys):
    """python -m glint_word2vec_b200 train / synonyms / export."""
    import json as _json
    from glint_word2vec_b200.cli import main
    from glint_word2vec_b200.data.synthetic import synthetic_capitals_corpus
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("\n".join(" ".join(s) for s in synthetic_capitals_corpus(n_sent=2500)), encoding="utf-8")
    out = str(tmp_path / "model")
    assert main(["train", str(corpus), "--out", out, "--vector-size", "32", "--step-size", "0.025", "--seed", "3",
                 "--config", "device=cpu", "--config", "subsample_mode=reference"]) == 0
    info = _json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert info["vector_size"] == 32 and info["words"] > 20 and os.path.isdir(os.path.join(out, "matrix"))
    assert main(["synonyms", out, "wien", "no-such-word", "--num", "3", "--config", "device=cpu"]) == 0
    lines = [_json.loads(l) for l in capsys.readouterr().out.strip().splitlines()]
    assert lines[0]["word"] == "wien" and len(lines[0]["synonyms"]) == 3
    assert lines[1] == {"word": "no-such-word", "error": "not in vocabulary"}
    local = str(tmp_path / "local")
    assert main(["export", out, local, "--config", "device=cpu"]) == 0
    assert os.path.isdir(os.path.join(local, "data"))


def test_parameter_server_config_reaches_the_engine_options():
    """hot_row_cap / sampler / step_tokens of parameterServerConfig must arrive in EngineOptions (a filter list once
    dropped hot_row_cap and sampler silently)."""
    from glint_word2vec_b200.data.synthetic import synthetic_capitals_corpus
    est = ServerSideGlintWord2Vec(inputCol="s", outputCol="v", vectorSize=8, minCount=5, seed=1, numParameterServers=1, numPartitions=3,
                                  unigramTableSize=1000000,
                                  parameterServerConfig={"device": "cpu", "hot_row_cap": 7.5, "sampler": "table",
                                                         "step_tokens": 4096, "neg_sharing": "tile", "tile_negatives": 32})
    m = est.fit(synthetic_capitals_corpus()[:300])
    try:
        eng = m._require_handle().engine
        assert eng.opts.hot_row_cap == 7.5 and eng.opts.sampler == "table" and eng.opts.step_tokens == 4096
        assert eng.opts.num_partitions == 3 and eng.opts.unigram_table_size == 1000000
        assert eng.cfg.neg_sharing == "tile" and eng.cfg.tile_negatives == 32
    finally:
        m.stop()
