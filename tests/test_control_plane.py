"""Control plane of the shard-server group: safe wire format, authentication, op whitelist, path confinement and the
cross-rank agreement step (pickle RPC behind a constant key; a handler failing on one rank wedged the
group)."""
import json
import os
import socket
import subprocess
import sys
import threading
import time

import numpy as np
import pytest

from glint_word2vec_b200.parallel import cluster, server as srv, wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pair():
    a, b = socket.socketpair()
    return a, b


def test_wire_roundtrip_arrays_and_scalars():
    a, b = _pair()
    obj = {"op": "x", "args": [np.arange(7, dtype=np.int64), 3, 2.5, "s", None, True,
                               {"m": np.ones((2, 3), np.float32), "empty": np.zeros((0, 4), np.int32)}], "kwargs": {}}
    t = threading.Thread(target=wire.send_msg, args=(a, obj))
    t.start()
    got = wire.recv_msg(b)
    t.join()
    assert got["op"] == "x" and got["args"][1:6] == [3, 2.5, "s", None, True]
    assert np.array_equal(got["args"][0], np.arange(7)) and got["args"][0].dtype == np.int64
    assert got["args"][6]["m"].shape == (2, 3) and got["args"][6]["empty"].shape == (0, 4)


def test_wire_rejects_objects_and_bad_frames():
    a, b = _pair()
    with pytest.raises(wire.WireError):
        wire.send_msg(a, {"x": np.array([object()], dtype=object)})
    with pytest.raises(wire.WireError):
        wire.send_msg(a, {1: 2})
    # a frame whose buffer size does not match dtype/shape is refused before any allocation of that size
    import struct
    head = json.dumps({"body": {"__nd__": 0, "dtype": "float32", "shape": [4]}, "sizes": [1 << 30]}).encode()
    a.sendall(struct.pack("<I", len(head)) + head)
    with pytest.raises(wire.WireError):
        wire.recv_msg(b)


def _start_group(tmp_path, n=1, conf=None, env_extra=None):
    ready = str(tmp_path / "ready.json")
    port = srv.free_port()
    cmd = [sys.executable, "-m", "glint_word2vec_b200.parallel.server", "--num-servers", str(n), "--port", str(port),
           "--device", "cpu", "--ready-file", ready]
    if conf is not None:
        cp = str(tmp_path / "conf.json")
        with open(cp, "w") as f:
            json.dump(conf, f)
        cmd += ["-c", cp]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    env.pop("GW2V_SERVER_SECRET", None)
    env.update(env_extra or {})
    proc = subprocess.Popen(cmd, env=env)
    t0 = time.time()
    while not os.path.exists(ready):
        assert proc.poll() is None, "server group died"
        assert time.time() - t0 < 180
        time.sleep(0.1)
    with open(ready) as f:
        info = json.load(f)
    return proc, port, info, ready


def _stop(proc, port, secret):
    try:
        cluster.RemoteHandle("127.0.0.1", port, secret=secret)._call("shutdown", timeout=10)
    except Exception:
        pass
    try:
        proc.wait(timeout=20)
    except Exception:
        proc.kill()


def test_server_authentication_whitelist_and_paths(tmp_path):
    root = tmp_path / "data"
    root.mkdir()
    proc, port, info, ready = _start_group(tmp_path, 1, {"data_root": str(root)})
    secret = info["secret"].encode()
    try:
        assert oct(os.stat(ready).st_mode & 0o777) == "0o600"          # the generated secret is private to this user
        with pytest.raises((wire.AuthError, EOFError, ConnectionError, OSError)):
            cluster.RemoteHandle("127.0.0.1", port, secret=b"wrong")
        h = cluster.RemoteHandle("127.0.0.1", port, secret=secret)
        # only whitelisted ops exist: attribute names of the server object are not reachable
        for op in ("execute", "_path", "__class__", "eval"):
            with pytest.raises(cluster.ServerError, match="unknown op"):
                h._call(op)
        from glint_word2vec_b200.models.sgns import SGNSConfig
        h.create(SGNSConfig(50, 8), {"subsample_mode": "reference"}, np.arange(50, 0, -1))
        with pytest.raises(cluster.ServerError, match="outside the server's data_root"):
            h.save(str(tmp_path / "elsewhere"))
        with pytest.raises(cluster.ServerError, match="outside the server's data_root"):
            h.save(str(root / ".." / "escape"))
        h.save(str(root / "m1"))
        assert os.path.exists(root / "m1" / "matrix" / "meta.json")
        # raw garbage on the socket is dropped without harming the group
        s = socket.create_connection(("127.0.0.1", port))
        s.sendall(b"\x80\x04\x95garbage-pickle")
        s.close()
        assert h._call("info")["world"] == 1
    finally:
        _stop(proc, port, secret)


def test_non_loopback_bind_needs_a_secret(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    env.pop("GW2V_SERVER_SECRET", None)
    env.pop("GW2V_SERVER_SECRET_FILE", None)
    r = subprocess.run([sys.executable, "-m", "glint_word2vec_b200.parallel.server", "--num-servers", "1", "--port",
                        str(srv.free_port()), "--device", "cpu", "--bind", "0.0.0.0"], env=env, capture_output=True,
                       text=True, timeout=120)
    assert r.returncode != 0 and "without a configured secret" in (r.stderr + r.stdout)


def test_failure_on_one_rank_fails_the_request_everywhere_and_the_group_survives(tmp_path):
    """A handler that fails on ONE rank: the agreement step turns it into an error for the client, the half-built
    matrix is dropped on every rank and the group keeps serving (round 1: the other ranks waited 30 days)."""
    proc, port, info, _ = _start_group(tmp_path, 2, None, {"GW2V_TEST_FAIL_OP": "set_noise", "GW2V_TEST_FAIL_RANK": "1"})
    secret = info["secret"].encode()
    try:
        from glint_word2vec_b200.models.sgns import SGNSConfig
        h = cluster.RemoteHandle("127.0.0.1", port, secret=secret)
        h.create(SGNSConfig(60, 8), {"subsample_mode": "reference"}, np.arange(60, 0, -1))
        assert h.pull([1, 2]).shape == (2, 8)
        with pytest.raises(cluster.ServerError, match="failed on 1 of 2 shards"):
            h._call("set_noise", h.matrix_id, np.arange(60, 0, -1))
        # the matrix is gone on BOTH ranks (consistent state), and the group still answers
        with pytest.raises(cluster.ServerError, match="no matrix"):
            h.pull([1, 2])
        assert h._call("info")["matrices"] == []
        h2 = cluster.RemoteHandle("127.0.0.1", port, secret=secret)
        h2.create(SGNSConfig(30, 4), {"subsample_mode": "reference"}, np.arange(30, 0, -1))
        assert h2.norms().shape == (30,)
    finally:
        _stop(proc, port, secret)


def test_fit_text_file_streams_the_corpus_to_a_server_group(tmp_path):
    """fitTextFile in integrated server mode: the encoded corpus is streamed to disk and the two shard servers receive
    its PREFIX (nothing is pickled, no rank holds a private copy); the model equals the single-process one."""
    import numpy as np
    from glint_word2vec_b200 import ServerSideGlintWord2Vec
    from glint_word2vec_b200.data.synthetic import synthetic_capitals_corpus
    sentences = synthetic_capitals_corpus(seed=3)[:1200]
    path = tmp_path / "corpus.txt"
    path.write_text("\n".join(" ".join(s) for s in sentences), encoding="utf-8")
    kw = dict(inputCol="s", outputCol="v", vectorSize=16, minCount=2, seed=5)
    m1 = ServerSideGlintWord2Vec(numParameterServers=1, parameterServerConfig={"device": "cpu"}, **kw).fitTextFile(str(path))
    m2 = ServerSideGlintWord2Vec(numParameterServers=2,
                                 parameterServerConfig={"device": "cpu", "stream_threshold_bytes": 0,
                                                        "scratch_dir": str(tmp_path)}, **kw).fitTextFile(str(path))
    try:
        assert m2._require_handle().num_shards == 2
        v1, v2 = m1.getVectorsMap(), m2.getVectorsMap()
        assert v1.keys() == v2.keys()
        assert all(np.allclose(v1[w], v2[w], atol=1e-5) for w in list(v1)[:80])
        assert not [p for p in tmp_path.iterdir() if p.name.startswith("gw2v-corpus-")]      # the temporary stream is gone
    finally:
        m1.stop()
        m2.stop()
