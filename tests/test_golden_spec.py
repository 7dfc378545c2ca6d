"""Golden integration tier: the 15 scenarios of the reference's only test suite
(ServerSideGlintWord2VecSpec, SPEC:83-415) on the same corpus, hyper-parameters
and thresholds -- integrated shard group, separate shard-server group, load
variants, transform variants, synonyms, analogies, getVectors, toLocal.

Runs on the CPU path (2 Gloo shards = the spec's 2 parameter servers).  The GPU
variant of the quality gate is tests/test_gpu_ops.py::test_fit_on_gpu_golden.
"""
import json
import os
import pickle
import subprocess
import sys
import time

import numpy as np
import pandas as pd
import pytest

from glint_word2vec_b200 import (MLlibServerSideGlintWord2VecModel, ServerSideGlintWord2Vec,
                                 ServerSideGlintWord2VecModel, Word2VecModel)
from glint_word2vec_b200.parallel import server as srv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PS_CONFIG = {"device": "cpu", "subsample_mode": "reference"}          # Q1: the reference's sub-sampling is inert


def _estimator(**kw):
    # SPEC:87-94: seed 1, stepSize 0.025, 2 partitions, 2 parameter servers, unigram table 1e6
    args = dict(seed=1, stepSize=0.025, numPartitions=2, numParameterServers=2, inputCol="sentence",
                outputCol="model", unigramTableSize=1000000, parameterServerConfig=PS_CONFIG)
    args.update(kw)
    return ServerSideGlintWord2Vec(**args)


@pytest.fixture(scope="module")
def separate_cluster(tmp_path_factory):
    """A long-lived stand-alone shard-server group (the `glint.Main` application of SBT:50-59).

    With ``GW2V_IT_SERVER_HOST=ip:port`` (exported by ``scripts/it_env.sh exec`` / ``scripts/run_integration.py``)
    the suite attaches to that externally managed group instead of starting its own."""
    ext = os.environ.get("GW2V_IT_SERVER_HOST")
    if ext:
        yield ext, None
        return
    tmp = tmp_path_factory.mktemp("sep")
    ready = str(tmp / "ready.json")
    port = srv.free_port()
    conf = str(tmp / "separate.json")
    with open(conf, "w") as f:
        json.dump({"subsample_mode": "reference"}, f)                  # cf. separate-glint.conf (SEPCONF)
    # a separate group is reached with a per-deployment secret shared by server and clients (parallel/wire.py)
    os.environ["GW2V_SERVER_SECRET"] = "it-" + os.urandom(8).hex()
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    proc = subprocess.Popen([sys.executable, "-m", "glint_word2vec_b200.parallel.server", "--num-servers", "2",
                             "--port", str(port), "--bind", "127.0.0.1", "--device", "cpu", "-c", conf,
                             "--ready-file", ready], env=env)
    t0 = time.time()
    while not os.path.exists(ready):
        assert proc.poll() is None, "separate server group died"
        assert time.time() - t0 < 180
        time.sleep(0.2)
    yield f"127.0.0.1:{port}", proc
    if proc.poll() is None:
        try:
            from glint_word2vec_b200.parallel.cluster import RemoteHandle
            RemoteHandle("127.0.0.1", port)._call("shutdown", timeout=10)
        except Exception:
            pass
        try:
            proc.wait(timeout=20)
        except Exception:
            proc.kill()


@pytest.fixture(scope="module")
def trained(corpus_sentences, tmp_path_factory):
    """Scenario 1: train + save with an integrated shard group (SPEC:83-106)."""
    path = str(tmp_path_factory.mktemp("models") / "capitals.model")
    model = _estimator().fit(pd.DataFrame({"sentence": corpus_sentences}))
    try:
        model.save(path)
        assert os.path.isdir(path)
    finally:
        model.stop()
    return path


@pytest.fixture(scope="module")
def loaded(trained):
    m = ServerSideGlintWord2VecModel.load(trained)                     # scenario 3
    yield m
    m.stop()


def _check_params(m):
    # SPEC:146-153 / 167-174 / 187-193
    assert m.getSeed() == 1 and m.getNumPartitions() == 2 and m.getNumParameterServers() == 2
    assert m.getInputCol() == "sentence" and m.getOutputCol() == "model"
    assert m.getUnigramTableSize() == 1000000 and m.getVectorSize() == 100


def test_01_train_and_save_integrated(trained):
    assert sorted(os.listdir(trained)) == ["matrix", "metadata", "words"]
    meta = json.load(open(os.path.join(trained, "matrix", "meta.json")))
    assert meta["num_shards"] == 2 and meta["vocab_size"] == len(
        open(os.path.join(trained, "words", "part-00000"), encoding="utf-8").read().split("\n")) - 1


def test_02_train_and_save_on_separate_cluster(corpus_sentences, separate_cluster, tmp_path_factory):
    host, _ = separate_cluster
    path = str(tmp_path_factory.mktemp("sepmodels") / "capitals.model")
    model = _estimator(parameterServerHost=host).fit(pd.DataFrame({"sentence": corpus_sentences[:100]}))  # SPEC:111-112
    try:
        model.save(path)
        assert os.path.isdir(path)
        assert model.getParameterServerHost() == host
    finally:
        model.stop()                                                   # destroys the matrix, keeps the cluster
    pytest.sep_model_path = path


def test_03_load_integrated(loaded):
    _check_params(loaded)
    assert loaded.numWords == 3611 or not os.path.exists("/root/reference")


def test_04_load_onto_separate_cluster(trained, separate_cluster):
    host, _ = separate_cluster
    m = ServerSideGlintWord2VecModel.load(trained, host, {"subsample_mode": "reference"})     # SPEC:157-176
    try:
        _check_params(m)
        assert m.getParameterServerHost() == host
        assert len(m.findSynonymsArray("wien", 3)) == 3
        # a model attached to a server group is serialisable (usable in closures, SPEC:230,250)
        m2 = pickle.loads(pickle.dumps(m))
        assert np.allclose(m2.transformWord("wien"), m.transformWord("wien"))
    finally:
        m.stop()


def test_05_load_model_trained_on_separate_cluster_then_terminate(separate_cluster):
    host, proc = separate_cluster
    path = getattr(pytest, "sep_model_path", None)
    if path is None:
        pytest.skip("scenario 2 did not run")
    m = ServerSideGlintWord2VecModel.load(path)                        # host comes from the saved metadata
    try:
        assert m.getParameterServerHost() == host
        assert m.getSeed() == 1 and m.getVectorSize() == 100
    finally:
        m.stop(terminateOtherClients=True)                             # SPEC:194
    if proc is not None:
        proc.wait(timeout=30)
        assert proc.poll() is not None
    else:                                                              # externally managed group: the port must close
        import socket
        h, p = srv.parse_host(host)
        t0 = time.time()
        while True:
            try:
                socket.create_connection((h, p), timeout=1).close()
            except OSError:
                break
            assert time.time() - t0 < 30, "separate server group still accepting connections"
            time.sleep(0.5)


WORDS4 = ["österreich", "wien", "deutschland", "berlin"]


def test_06_transform_dataframe(loaded):
    out = loaded.transform(pd.DataFrame({"sentence": [[w] for w in WORDS4]}))
    vecs = out["model"].tolist()
    assert len(vecs) == 4 and all(len(v) == 100 and np.sum(v) != 0 for v in vecs)           # SPEC:198-217


def test_07_mllib_transform_word(loaded):
    ml = MLlibServerSideGlintWord2VecModel(loaded)
    vecs = [ml.transform(w) for w in WORDS4]
    assert all(len(v) == 100 and np.sum(v) != 0 for v in vecs)                              # SPEC:220-238


def test_08_mllib_transform_iterator(loaded):
    ml = MLlibServerSideGlintWord2VecModel(loaded)
    vecs = list(ml.transform(iter(WORDS4)))
    assert len(vecs) == 4 and all(len(v) == 100 and np.sum(v) != 0 for v in vecs)           # SPEC:240-258


def test_09_transform_keeps_other_columns(loaded):
    df = pd.DataFrame({"id": range(4), "sentence": [[w] for w in WORDS4], "label": list("abcd")})
    out = loaded.transform(df)
    assert list(out.columns) == ["id", "sentence", "label", "model"]                        # SPEC:260-288


def test_10_synonyms_array(loaded):
    syn = loaded.findSynonymsArray("österreich", 10)
    assert len(syn) == 10
    d = dict(syn)
    assert "wien" in d and d["wien"] > 0.9                                                  # SPEC:297-301


def test_11_synonyms_dataframe(loaded):
    df = loaded.findSynonyms("österreich", 10)
    assert list(df.columns) == ["word", "similarity"] and len(df) == 10
    row = df[df["word"] == "wien"]
    assert len(row) == 1 and float(row["similarity"].iloc[0]) > 0.9                         # SPEC:307-325


def _analogy_vector(model):
    out = model.transform(pd.DataFrame({"sentence": [["wien"], ["österreich"], ["deutschland"]]}))
    w, o, d = out["model"].tolist()
    return w - o + d


def test_12_analogy_array(loaded):
    syn = dict(loaded.findSynonymsArray(_analogy_vector(loaded), 10))
    assert "berlin" in syn and syn["berlin"] > 0.9                                          # SPEC:327-352


def test_13_analogy_dataframe(loaded):
    df = loaded.findSynonyms(_analogy_vector(loaded), 10)
    row = df[df["word"] == "berlin"]
    assert len(row) == 1 and float(row["similarity"].iloc[0]) > 0.9                         # SPEC:354-382


def test_14_get_vectors(loaded):
    gv = loaded.getVectors()
    assert list(gv.columns) == ["word", "vector"] and len(gv) == loaded.numWords            # SPEC:384-398


def test_15_to_local_and_save(loaded, tmp_path):
    local = loaded.toLocal()
    assert isinstance(local, Word2VecModel)
    p = str(tmp_path / "local.model")
    local.save(p)
    assert os.path.isdir(p)                                                                 # SPEC:400-415
    assert "wien" in dict(local.findSynonymsArray("österreich", 10))
