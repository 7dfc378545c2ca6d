"""Client-side handles to the sharded matrices (Glint ``Client`` +
``BigWord2VecMatrix`` + ``GranularBigWord2VecMatrix`` in one object).

Three ways to reach the shards, mirroring the reference's bootstrap (C9,
MLLIB:351-362 and the loader MLLIB:716-721):

``InProcessHandle``
    The calling process IS a shard (world size 1, or one rank of an SPMD job
    launched with ``torchrun`` where every rank runs the same user program and
    the engine collectives line up).  No RPC at all.

``spawn_integrated`` -> ``RemoteHandle``
    "integrated" mode: ``parameterServerHost == ""``.  The client starts a
    server group (one process per GPU) as children of this process, like
    ``Client.runWithWord2VecMatrixOnSpark`` occupying executors, and talks to
    rank 0 over TCP.  ``terminate()`` stops it.

``RemoteHandle(host)``
    "separate" mode: ``parameterServerHost`` names a running server group
    started with ``python -m glint_word2vec_b200.parallel.server``.

There is no message-size cap here, so ``GranularBigWord2VecMatrix``'s 10 000
element chunking (MLLIB:83-85,362) has no counterpart.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile
import time
import uuid
import socket
from typing import Optional, Tuple

import numpy as np
import torch

from ..data.corpus import EncodedCorpus
from ..models import matrix_io, trainer
from ..models.engine import EngineOptions, ShardEngine
from ..models.sgns import SGNSConfig
from . import server as _server
from . import wire
from .comm import Comm, comm_from_env


def run_training(engine, corpus, lr, iters, train_words, metrics_path=None, train_opts=None):
    """Plain run, checkpointed run, or resume-from-checkpoint (models/checkpoint.py)."""
    from ..models import checkpoint
    o = dict(train_opts or {})
    ckdir, every = o.get("checkpoint_dir"), int(o.get("checkpoint_every_steps", 0) or 0)
    if ckdir and o.get("resume") and checkpoint.latest(ckdir):
        import json
        with open(os.path.join(checkpoint.latest(ckdir), "state.json")) as f:
            st = json.load(f)
        fp = checkpoint.fingerprint(engine, corpus)
        checkpoint.check_fingerprint(st, fp)                  # a stale directory of another run is an error, not a resume
        engine.opts.step_tokens = int(st["step_tokens"])
        from ..models import matrix_io as _mio
        loaded = _mio.load_matrix(checkpoint.latest(ckdir), engine.comm, engine.device, engine.opts)
        engine.syn0, engine.syn1 = loaded.syn0, loaded.syn1
        ck = checkpoint.Checkpointer(engine, ckdir, every, {k: st[k] for k in (
            "learning_rate", "num_iterations", "train_words", "step_tokens")}, fp=fp,
            run_id=st.get("run_id")) if every > 0 else None
        return trainer.train(engine, corpus, st["learning_rate"], st["num_iterations"], st["train_words"],
                             metrics_path=metrics_path, checkpoint_fn=ck, start_iteration=st["iteration"],
                             start_step=st["next_step"])
    if ckdir and every > 0:
        return checkpoint.train_with_checkpoints(engine, corpus, lr, iters, train_words, ckdir, every, metrics_path,
                                                 overwrite=bool(o.get("checkpoint_overwrite", False)))
    return trainer.train(engine, corpus, lr, iters, train_words, metrics_path=metrics_path)


class MatrixHandle:
    """Interface shared by all handles (the reference's call-site contract)."""
    num_shards: int = 1
    cols: int = 0                      # == vectorSize (``matrix.cols`` MLLIB:473)
    host: str = ""                     # what to persist as ``parameterServerHost``

    def fit(self, corpus: EncodedCorpus, lr: float, iters: int, train_words: int, metrics_path=None,
            train_opts=None) -> dict:
        raise NotImplementedError

    def pull(self, rows) -> np.ndarray: raise NotImplementedError
    def pull_average(self, rows_flat, offsets) -> np.ndarray: raise NotImplementedError
    def norms(self) -> np.ndarray: raise NotImplementedError
    def multiply(self, q) -> np.ndarray: raise NotImplementedError
    def top_k(self, queries, k) -> Tuple[np.ndarray, np.ndarray]: raise NotImplementedError
    def save(self, path: str, extra=None): raise NotImplementedError
    def destroy(self): raise NotImplementedError
    def terminate(self, terminate_other_clients: bool = False): raise NotImplementedError


# --------------------------------------------------------------------------- in process

class InProcessHandle(MatrixHandle):
    def __init__(self, engine: ShardEngine):
        self.engine = engine
        self.num_shards = engine.comm.world
        self.cols = engine.cfg.vector_size
        self.last_report: Optional[dict] = None

    @classmethod
    def create(cls, cfg: SGNSConfig, opts: dict, counts, device=None, comm: Optional[Comm] = None):
        eng = ShardEngine(cfg, comm=comm or comm_from_env(), device=device,
                          options=EngineOptions.from_dict(opts))
        eng.init_weights()
        eng.set_noise(np.asarray(counts))
        return cls(eng)

    @classmethod
    def load(cls, path: str, opts: dict, device=None, comm: Optional[Comm] = None):
        eng = matrix_io.load_matrix(path, comm or comm_from_env(), device, EngineOptions.from_dict(opts))
        return cls(eng)

    def fit(self, corpus, lr, iters, train_words, metrics_path=None, train_opts=None):
        rep = run_training(self.engine, corpus, lr, iters, train_words, metrics_path, train_opts)
        self.last_report = {k: getattr(rep, k) for k in
                            ("iterations", "steps", "words", "pairs", "loss_per_pair", "max_abs_dot",
                             "seconds", "final_alpha", "device_ms")}
        self.last_report["history"] = rep.history[-50:]
        return self.last_report

    def pull(self, rows):
        return self.engine.pull(np.asarray(rows, np.int64)).cpu().numpy()

    def pull_average(self, rows_flat, offsets):
        return self.engine.pull_average(np.asarray(rows_flat, np.int64),
                                        np.asarray(offsets, np.int64)).cpu().numpy()

    def norms(self):
        return self.engine.norms().cpu().numpy()

    def multiply(self, q):
        return self.engine.multiply(np.asarray(q, np.float32)).cpu().numpy()

    def top_k(self, queries, k):
        idx, sim = self.engine.top_k(np.asarray(queries, np.float32), int(k))
        return idx.numpy(), sim.numpy()

    def save(self, path, extra=None):
        matrix_io.save_matrix(self.engine, path, extra)

    def destroy(self):
        self.engine.destroy()

    def terminate(self, terminate_other_clients=False):
        pass


# --------------------------------------------------------------------------- remote

class ServerError(RuntimeError):
    pass


class RemoteHandle(MatrixHandle):
    """Talks to rank 0 of a shard-server group.  One short-lived connection per
    request, so several clients can share a group (cf. several Spark apps on
    one Glint cluster)."""

    def __init__(self, host: str, port: int, matrix_id: Optional[str] = None, owned_procs=None,
                 persist_host: str = "", secret: Optional[bytes] = None, scratch_dir: Optional[str] = None):
        self.addr = (host, port)
        self.matrix_id = matrix_id or uuid.uuid4().hex[:12]
        self._procs = owned_procs or []
        self.host = persist_host
        self._secret = secret if secret is not None else wire.load_secret()
        if self._secret is None:
            raise PermissionError("no server secret: set GW2V_SERVER_SECRET (or GW2V_SERVER_SECRET_FILE) to the secret "
                                  "of the shard-server group")
        self._scratch = scratch_dir
        info = self._call("info")
        self.num_shards = info["world"]
        self.device = info["device"]
        self.last_report: Optional[dict] = None

    # -- transport: JSON header + raw numpy buffers over an HMAC-authenticated socket (parallel/wire.py)
    def _call(self, op, *args, timeout: float = 300.0, **kwargs):
        deadline = time.time() + 30.0
        while True:
            try:
                conn = socket.create_connection(self.addr, timeout=30.0)
                break
            except (ConnectionRefusedError, OSError):
                if time.time() > deadline:
                    raise
                time.sleep(0.1)
        try:
            wire.client_handshake(conn, self._secret)
            wire.send_msg(conn, {"op": op, "args": list(args), "kwargs": kwargs})
            # the reference awaits RPCs with 1-5 minute time-outs (MLLIB:429,486,497)
            conn.settimeout(timeout)
            try:
                resp = wire.recv_msg(conn)
            except socket.timeout:
                raise TimeoutError(f"server did not answer {op!r} within {timeout}s") from None
        finally:
            conn.close()
        if not resp["ok"]:
            raise ServerError(resp["error"] + "\n" + resp.get("trace", ""))
        return resp["result"]

    # -- matrix lifecycle
    def create(self, cfg: SGNSConfig, opts: dict, counts):
        r = self._call("create", self.matrix_id, cfg.to_dict(), opts, np.asarray(counts, np.int64))
        self.cols = r["cols"]
        return self

    def load(self, path: str, opts: dict):
        r = self._call("load", self.matrix_id, path, opts, timeout=3600.0)
        self.cols = r["cols"]
        return r

    def fit(self, corpus, lr, iters, train_words, metrics_path=None, train_opts=None):
        """The corpus goes to the shards as a PATH: a disk-backed corpus is used as is, an in-memory one is written
        once to a scratch prefix; every rank memory-maps the same files (nothing is pickled or re-broadcast)."""
        tmp = None
        if corpus.prefix is None:
            tmp = tempfile.mkdtemp(prefix="gw2v_corpus_", dir=self._scratch)
            corpus = corpus.save(os.path.join(tmp, "corpus"))
        try:
            self.last_report = self._call("fit", self.matrix_id, {"prefix": corpus.prefix}, lr, iters, train_words,
                                          metrics_path, train_opts, timeout=7 * 24 * 3600.0)
        finally:
            if tmp is not None:
                import shutil
                shutil.rmtree(tmp, ignore_errors=True)
        return self.last_report

    def pull(self, rows):
        return self._call("pull", self.matrix_id, np.asarray(rows, np.int64))

    def pull_average(self, rows_flat, offsets):
        return self._call("pull_average", self.matrix_id, np.asarray(rows_flat, np.int64),
                          np.asarray(offsets, np.int64))

    def norms(self):
        return self._call("norms", self.matrix_id)

    def multiply(self, q):
        return self._call("multiply", self.matrix_id, np.asarray(q, np.float32))

    def top_k(self, queries, k):
        idx, sim = self._call("top_k", self.matrix_id, np.asarray(queries, np.float32), int(k))
        return idx, sim

    def save(self, path, extra=None):
        return self._call("save", self.matrix_id, path, extra, timeout=3600.0)

    def destroy(self):
        try:
            self._call("destroy", self.matrix_id)
        except (ConnectionError, OSError, ServerError):
            pass

    def terminate(self, terminate_other_clients: bool = False):
        """``client.terminateOnSpark(sc, terminateOtherClients)`` (MLLIB:666):
        an integrated group we spawned is always stopped; a separate group only
        when ``terminate_other_clients`` is true (SPEC:194)."""
        if self._procs or terminate_other_clients:
            try:
                self._call("shutdown", timeout=30.0)
            except Exception:
                pass
        for p in self._procs:
            try:
                p.wait(timeout=20)
            except Exception:
                p.kill()
        self._procs = []


def spawn_integrated(num_servers: int, device_type: str, options: Optional[dict] = None,
                     startup_timeout: float = 300.0) -> RemoteHandle:
    """Start a private server group (integrated mode) and connect to it."""
    port = _server.free_port()
    tmp = tempfile.mkdtemp(prefix="gw2v_srv_")
    ready = os.path.join(tmp, "ready.json")
    cmd = [sys.executable, "-m", "glint_word2vec_b200.parallel.server", "--num-servers", str(num_servers),
           "--port", str(port), "--bind", "127.0.0.1", "--device", device_type, "--ready-file", ready]
    if options:
        cfgp = os.path.join(tmp, "options.json")
        with open(cfgp, "w") as f:
            json.dump(options, f)
        cmd += ["-c", cfgp]
    env = dict(os.environ)
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_PORT"):
        env.pop(k, None)
    proc = subprocess.Popen(cmd, env=env)
    deadline = time.time() + startup_timeout
    while not os.path.exists(ready):
        if proc.poll() is not None:
            raise RuntimeError(f"shard-server group exited with code {proc.returncode} during start-up")
        if time.time() > deadline:
            proc.kill()
            raise TimeoutError("shard-server group did not come up")
        time.sleep(0.1)
    with open(ready) as f:
        info = json.load(f)
    # the group generated its own secret and wrote it into the 0600 ready-file: only this user can reach it
    secret = info["secret"].encode() if info.get("secret") else wire.load_secret(options)
    return RemoteHandle("127.0.0.1", port, owned_procs=[proc], persist_host="", secret=secret, scratch_dir=tmp)


def connect_separate(host: str) -> RemoteHandle:
    h, p = _server.parse_host(host)
    return RemoteHandle(h, p, persist_host=host)


def resolve_integrated_shards(num_parameter_servers: int) -> Tuple[int, str]:
    """How many column shards an integrated run really gets, and on what.

    ``numParameterServers`` is a request: on a GPU box it is capped by the number
    of visible GPUs (one shard per GPU); on a CPU-only box it is honoured as
    given (Gloo processes)."""
    if torch.cuda.is_available():
        return max(1, min(num_parameter_servers, torch.cuda.device_count())), "cuda"
    return max(1, num_parameter_servers), "cpu"
