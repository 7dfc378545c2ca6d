"""Shard-server group: the B200 counterpart of the Glint parameter-server cluster.

Reference: parameter servers either run inside the Spark application
("integrated", ``Client.runWithWord2VecMatrixOnSpark`` MLLIB:355) or as a
separate long-lived application started with
``spark-submit --class glint.Main ... spark [-c conf]`` whose master IP is then
passed as ``parameterServerHost`` (README.md:52-57, MLLIB:358-360).

Here a server group is S processes (one per GPU / column shard).  Rank 0 is the
"master": it listens on a TCP port, receives requests from clients, broadcasts
each request to the other ranks over the Gloo control group and all ranks
execute it in lock-step on their shard; rank 0 returns the result.  Several
matrices (models) can live on one group, like several ``BigWord2VecMatrix`` on
one Glint cluster.

Stand-alone ("separate") mode::

    python -m glint_word2vec_b200.parallel.server --num-servers 8 --port 13380 [-c conf.json]

prints ``master = <ip>:<port>`` -- pass it as ``parameterServerHost``.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import socket
import sys
import time
import traceback
import select
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist

from ..data.corpus import EncodedCorpus
from ..models import matrix_io
from ..models.engine import EngineOptions, ShardEngine
from ..models.sgns import SGNSConfig
from . import wire
from .comm import Comm, TorchDistComm, init_process_group

log = logging.getLogger("glint_word2vec_b200.server")

DEFAULT_PORT = 13370           # cf. glint.master.port 13380 in separate-glint.conf (SEPCONF:3)
HEARTBEAT_S = 20.0             # rank 0 broadcasts an idle beat at least this often, so collectives can time out
COLLECTIVE_TIMEOUT_S = 900.0   # a rank that waits longer than this inside a control collective gives up (group dies)
# the wire protocol: nothing else can be invoked, whatever the client sends
ALLOWED_OPS = frozenset({"info", "create", "fit", "pull", "pull_average", "norms", "multiply", "top_k", "save", "load",
                         "set_noise", "destroy"})
MUTATING_OPS = frozenset({"create", "fit", "load", "set_noise"})


def parse_host(host: str, default_port: int = DEFAULT_PORT):
    """``"ip"`` or ``"ip:port"`` -> (ip, port)."""
    if ":" in host:
        h, p = host.rsplit(":", 1)
        return h, int(p)
    return host, default_port


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class ShardServer:
    """Per-rank request executor.  Every handler runs on ALL ranks."""

    def __init__(self, comm: Comm, device: torch.device, default_options: Optional[dict] = None):
        self.comm = comm
        self.device = device
        self.default_options = dict(default_options or {})
        # when set, every path a client names (save / load / corpus / metrics / checkpoints) must resolve below it
        self.data_root = self.default_options.pop("data_root", None)
        for k in ("secret", "secret_file", "port", "bind"):
            self.default_options.pop(k, None)
        self.engines: Dict[str, ShardEngine] = {}
        self.reports: Dict[str, dict] = {}

    # -- helpers
    def _opts(self, opts: Optional[dict]) -> EngineOptions:
        merged = dict(self.default_options)
        merged.update(opts or {})
        return EngineOptions.from_dict(merged)

    def _path(self, path: Optional[str]) -> Optional[str]:
        """Validate a client-supplied filesystem path against ``data_root``."""
        if path is None:
            return None
        if not isinstance(path, str) or "\x00" in path:
            raise ValueError("bad path")
        if self.data_root:
            root = os.path.realpath(self.data_root)
            real = os.path.realpath(path)
            if real != root and not real.startswith(root + os.sep):
                raise PermissionError(f"path {path!r} is outside the server's data_root")
        return path

    def _eng(self, mid: str) -> ShardEngine:
        if mid not in self.engines:
            raise KeyError(f"no matrix {mid!r} on this server group (destroyed or never created)")
        return self.engines[mid]

    # -- handlers (names are the wire protocol)
    def op_info(self):
        return {"world": self.comm.world, "device": str(self.device),
                "matrices": sorted(self.engines), "pid": os.getpid()}

    def op_create(self, mid, cfg, opts, counts):
        eng = ShardEngine(SGNSConfig(**cfg), comm=self.comm, device=self.device, options=self._opts(opts))
        eng.init_weights()
        eng.set_noise(np.asarray(counts))
        self.engines[mid] = eng
        return {"cols": eng.cfg.vector_size, "shards": self.comm.world}

    def op_fit(self, mid, corpus, lr, iters, train_words, metrics_path=None, train_opts=None):
        """``corpus``: ``{"prefix": p}`` -- every rank memory-maps ``p.tokens.i32`` / ``p.offsets.i64`` (the corpus
        never travels through the control plane, cf. RDD partitions MLLIB:335-345) -- or ``{"tokens": .., "offsets": ..}``
        for small in-memory corpora."""
        from .cluster import run_training
        eng = self._eng(mid)
        if "prefix" in corpus:
            enc = EncodedCorpus.open(self._path(corpus["prefix"]))
        else:
            enc = EncodedCorpus(np.asarray(corpus["tokens"], np.int32), np.asarray(corpus["offsets"], np.int64))
        train_opts = dict(train_opts or {})
        if train_opts.get("checkpoint_dir"):
            self._path(train_opts["checkpoint_dir"])
        rep = run_training(eng, enc, float(lr), int(iters), int(train_words), self._path(metrics_path), train_opts)
        out = {k: getattr(rep, k) for k in ("iterations", "steps", "words", "pairs", "loss_per_pair",
                                             "max_abs_dot", "seconds", "final_alpha", "device_ms")}
        out["history"] = rep.history[-50:]
        self.reports[mid] = out
        return out

    def op_pull(self, mid, rows):
        return self._eng(mid).pull(np.asarray(rows, np.int64)).cpu().numpy()

    def op_pull_average(self, mid, rows_flat, offsets):
        return self._eng(mid).pull_average(np.asarray(rows_flat, np.int64),
                                           np.asarray(offsets, np.int64)).cpu().numpy()

    def op_norms(self, mid):
        return self._eng(mid).norms().cpu().numpy()

    def op_multiply(self, mid, q):
        return self._eng(mid).multiply(np.asarray(q, np.float32)).cpu().numpy()

    def op_top_k(self, mid, queries, k):
        idx, sim = self._eng(mid).top_k(np.asarray(queries, np.float32), int(k))
        return idx.numpy(), sim.numpy()

    def op_save(self, mid, path, extra=None):
        matrix_io.save_matrix(self._eng(mid), self._path(path), extra)
        return True

    def op_load(self, mid, path, opts=None, with_syn1=True):
        eng = matrix_io.load_matrix(self._path(path), self.comm, self.device, self._opts(opts), with_syn1)
        self.engines[mid] = eng
        return {"cols": eng.cfg.vector_size, "shards": self.comm.world,
                "vocab_size": eng.cfg.vocab_size, "config": eng.cfg.to_dict()}

    def op_set_noise(self, mid, counts):
        self._eng(mid).set_noise(np.asarray(counts))
        return True

    def op_destroy(self, mid):
        eng = self.engines.pop(mid, None)
        if eng is not None:
            eng.destroy()
        return True

    def execute(self, req: dict):
        op = req.get("op")
        if op not in ALLOWED_OPS:
            raise ValueError(f"unknown op {op!r}")
        args, kwargs = req.get("args", []), req.get("kwargs", {})
        if not isinstance(args, (list, tuple)) or not isinstance(kwargs, dict):
            raise ValueError("malformed request")
        # fault injection for the tests of the agreement step: fail `op` on one rank only
        if os.environ.get("GW2V_TEST_FAIL_OP") == op and os.environ.get("GW2V_TEST_FAIL_RANK") == str(self.comm.rank):
            raise RuntimeError("injected failure")
        return getattr(self, "op_" + op)(*args, **kwargs)


def _bcast_request(comm: Comm, req):
    """Rank 0 -> all ranks on the Gloo control group.  The object was decoded from the safe wire format on rank 0
    (plain dict / list / numpy), so what the ranks exchange among themselves is only ever data this group built."""
    if comm.world == 1:
        return req
    return comm.broadcast_object(req, src=0)


def _is_loopback(bind: str) -> bool:
    return bind in ("127.0.0.1", "localhost", "::1")


def serve(comm: Comm, device: torch.device, port: int, bind: str = "127.0.0.1",
          options: Optional[dict] = None, ready_file: Optional[str] = None):
    """Run the request loop until a ``shutdown`` request arrives."""
    configured = wire.load_secret(options)
    server = ShardServer(comm, device, options)
    srv = None
    secret = configured
    if comm.rank == 0:
        if secret is None:
            if not _is_loopback(bind):
                raise SystemExit("refusing to listen on %s without a configured secret: set GW2V_SERVER_SECRET, or "
                                 "`secret` / `secret_file` in the server config" % bind)
            secret = wire.new_secret()          # private to whoever can read the 0600 ready-file (the spawning client)
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((bind, port))
        srv.listen(32)
        ip = "127.0.0.1" if _is_loopback(bind) else _local_ip()
        # the reference prints the master IP to the log for the user to copy (README.md:56)
        print(f"master = {ip}:{port}", flush=True)
        if ready_file:
            fd = os.open(ready_file + ".tmp", os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
            with os.fdopen(fd, "w") as f:
                json.dump({"host": ip, "port": port, "pid": os.getpid(), "world": comm.world,
                           "secret": None if configured else secret.decode()}, f)
            os.replace(ready_file + ".tmp", ready_file)
    conn = None
    running = True
    while running:
        req = None
        if comm.rank == 0:
            req = {"op": "_idle"}
            try:
                if conn is None:
                    r, _, _ = select.select([srv], [], [], HEARTBEAT_S)
                    if r:
                        c, _addr = srv.accept()
                        try:
                            c.settimeout(30.0)
                            wire.server_handshake(c, secret)
                            c.settimeout(None)
                            conn = c
                        except Exception as e:
                            log.warning("rejected connection: %s", e)
                            c.close()
                if conn is not None:
                    r, _, _ = select.select([conn], [], [], HEARTBEAT_S)
                    if r:
                        conn.settimeout(600.0)
                        msg = wire.recv_msg(conn)
                        conn.settimeout(None)
                        if not isinstance(msg, dict) or not isinstance(msg.get("op"), str):
                            raise wire.WireError("malformed request")
                        req = msg
            except (EOFError, ConnectionError, OSError, wire.WireError, ValueError) as e:
                if not isinstance(e, EOFError):
                    log.warning("dropping connection: %s", e)
                if conn is not None:
                    conn.close()
                    conn = None
        req = _bcast_request(comm, req)
        op = req["op"]
        if op == "_idle":                       # heartbeat: keeps every rank inside short-lived collectives only
            continue
        if op == "shutdown":
            running = False
            resp = {"ok": True, "result": True}
        else:
            try:
                result = server.execute(req)
                resp = {"ok": True, "result": result}
            except Exception as e:
                resp = {"ok": False, "error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()}
            resp = _agree(comm, server, req, resp)
        if comm.rank == 0 and conn is not None:
            try:
                wire.send_msg(conn, resp)
            except (ConnectionError, OSError, wire.WireError) as e:
                log.warning("could not answer %s: %s", op, e)
                conn.close()
                conn = None
    if conn is not None:
        conn.close()
    if srv is not None:
        srv.close()
    for mid in list(server.engines):
        server.op_destroy(mid)


def _agree(comm: Comm, server: "ShardServer", req: dict, resp: dict) -> dict:
    """Every rank reports whether its handler succeeded; if they disagree (a shard file unreadable on one rank, an
    OOM on one GPU, ...) the request fails on ALL ranks, and a matrix that a mutating op left half-built is dropped
    everywhere so that the group stays consistent and keeps serving."""
    if comm.world == 1:
        return resp
    status = comm.gather_objects((bool(resp["ok"]), resp.get("error")), dst=0)
    verdict = None
    if comm.rank == 0:
        bad = [(r, err) for r, (ok, err) in enumerate(status) if not ok]
        verdict = {"bad": [[r, err] for r, err in bad], "mixed": 0 < len(bad) < comm.world}
    verdict = comm.broadcast_object(verdict, src=0)
    if verdict["bad"]:
        args = req.get("args") or [None]
        if req.get("op") in MUTATING_OPS and isinstance(args[0], str) and (verdict["mixed"] or req["op"] != "fit"):
            server.op_destroy(args[0])
        if resp["ok"] or verdict["mixed"]:
            ranks = ", ".join(f"rank {r}: {err}" for r, err in verdict["bad"])
            resp = {"ok": False, "error": f"request {req.get('op')!r} failed on {len(verdict['bad'])} of {comm.world} "
                                           f"shards ({ranks})", "trace": resp.get("trace", "")}
    return resp


def _local_ip() -> str:
    try:
        s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        s.connect(("10.255.255.255", 1))
        ip = s.getsockname()[0]
        s.close()
        return ip
    except Exception:
        return "127.0.0.1"


def rank_main(rank: int, world: int, port: int, master_port: int, device_type: str,
              options: Optional[dict], ready_file: Optional[str], bind: str = "127.0.0.1"):
    """Entry point of one server rank (one process per GPU)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(master_port)
    os.environ["RANK"] = str(rank)
    os.environ["WORLD_SIZE"] = str(world)
    os.environ["LOCAL_RANK"] = str(rank)
    if device_type == "cuda":
        torch.cuda.set_device(rank)
        device = torch.device("cuda", rank)
    else:
        device = torch.device("cpu")
        torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    comm: Comm
    if world > 1:
        # rank 0 sends an idle beat every HEARTBEAT_S, so no rank ever waits long inside a control collective unless a
        # peer is stuck or dead: then the collective times out, the rank exits and the group monitor stops the group
        init_process_group(rank=rank, world=world, master_port=master_port, device=device,
                           timeout_s=float((options or {}).get("collective_timeout_s", COLLECTIVE_TIMEOUT_S)))
        comm = TorchDistComm()
    else:
        comm = Comm()
    try:
        serve(comm, device, port, bind=bind, options=options, ready_file=ready_file)
    finally:
        if world > 1 and dist.is_initialized():
            try:
                dist.destroy_process_group()
            except Exception:
                pass


def main(argv=None):
    ap = argparse.ArgumentParser(description="stand-alone shard-server group (cf. glint.Main)")
    ap.add_argument("--num-servers", "-n", type=int, default=0, help="column shards (0 = all GPUs, or 1 on CPU)")
    ap.add_argument("--port", type=int, default=DEFAULT_PORT)
    ap.add_argument("--bind", default="127.0.0.1",
                    help="listen address; anything but loopback requires a configured secret (GW2V_SERVER_SECRET, "
                         "or `secret` / `secret_file` in the -c config)")
    ap.add_argument("--device", default="auto", choices=["auto", "cuda", "cpu"])
    ap.add_argument("-c", "--config", default=None, help="JSON file with engine options (cf. `-c separate-glint.conf`)")
    ap.add_argument("--ready-file", default=None)
    ap.add_argument("--rank", type=int, default=None, help=argparse.SUPPRESS)
    ap.add_argument("--master-port", type=int, default=None, help=argparse.SUPPRESS)
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    device_type = args.device
    if device_type == "auto":
        device_type = "cuda" if torch.cuda.is_available() else "cpu"
    world = args.num_servers
    if world <= 0:
        world = torch.cuda.device_count() if device_type == "cuda" else 1
    options = None
    if args.config:
        with open(args.config) as f:
            options = json.load(f)
        port = int(options.get("port", args.port)) if args.port == DEFAULT_PORT else args.port
    else:
        port = args.port
    if args.rank is not None:                      # child rank
        rank_main(args.rank, world, port, args.master_port, device_type, options, args.ready_file, args.bind)
        return
    master_port = free_port()
    if world == 1:
        rank_main(0, 1, port, master_port, device_type, options, args.ready_file, args.bind)
        return
    import subprocess
    procs = []
    for r in range(world):
        cmd = [sys.executable, "-m", "glint_word2vec_b200.parallel.server", "--num-servers", str(world),
               "--port", str(port), "--bind", args.bind, "--device", device_type, "--rank", str(r),
               "--master-port", str(master_port)]
        if args.config:
            cmd += ["-c", args.config]
        if args.ready_file:
            cmd += ["--ready-file", args.ready_file]
        procs.append(subprocess.Popen(cmd))
    # Failure detection: losing a shard loses its column slice, so the group cannot continue.
    # The first rank that exits abnormally takes the whole group down (exact pids of our own children only);
    # clients see their pending request fail and recover from the last checkpoint with a fresh group.
    rc = 0
    alive = list(procs)
    while alive:
        time.sleep(0.2)
        for p in list(alive):
            code = p.poll()
            if code is None:
                continue
            alive.remove(p)
            if code != 0:
                rc = rc or code
                log.error("shard process %d exited with code %s: stopping the group", p.pid, code)
                for q in alive:
                    q.terminate()
                deadline = time.time() + 10
                for q in alive:
                    try:
                        q.wait(timeout=max(0.1, deadline - time.time()))
                    except Exception:
                        q.kill()
                alive = []
                break
    sys.exit(rc)


if __name__ == "__main__":
    main()
