"""Wire format and authentication of the shard-server control plane.

The reference talks to its parameter servers with Akka remoting (SEPCONF:6,10); round 1 of this framework used
``multiprocessing.connection`` whose ``recv()`` is ``pickle.loads`` behind a constant auth key -- anybody who could
reach the port could run code on rank 0.  This module replaces it:

* frames are ``[u32 header length][JSON header][raw buffers]``: the header is plain JSON (dict / list / str / number /
  bool / None), numpy arrays travel as raw bytes referenced from the header as ``{"__nd__": i, "dtype": .., "shape": ..}``
  with a whitelist of dtypes.  Nothing on the wire is ever unpickled;
* every connection starts with an HMAC-SHA256 challenge/response over a per-deployment secret (``GW2V_SERVER_SECRET``,
  a ``secret`` / ``secret_file`` entry of the server config, or a random secret the server writes into its 0600
  ready-file for the client that spawned it).
"""
from __future__ import annotations

import hashlib
import hmac
import json
import os
import secrets
import socket
import struct
from typing import Any, List, Optional, Tuple

import numpy as np

MAX_HEADER = 64 << 20
MAX_BUFFER = 1 << 40
_DTYPES = {"int8", "uint8", "int16", "int32", "int64", "uint32", "uint64", "float16", "float32", "float64", "bool"}
MAGIC = b"GW2V1\n"


class WireError(RuntimeError):
    pass


class AuthError(WireError):
    pass


def _encode(obj: Any, bufs: List[np.ndarray]):
    if isinstance(obj, np.ndarray):
        if obj.dtype.name not in _DTYPES:
            raise WireError(f"dtype {obj.dtype} is not allowed on the wire")
        a = np.ascontiguousarray(obj)
        bufs.append(a)
        return {"__nd__": len(bufs) - 1, "dtype": a.dtype.name, "shape": list(a.shape)}
    if isinstance(obj, np.generic):
        return obj.item()
    if isinstance(obj, (list, tuple)):
        return [_encode(x, bufs) for x in obj]
    if isinstance(obj, dict):
        out = {}
        for k, v in obj.items():
            if not isinstance(k, str):
                raise WireError("only string keys are allowed on the wire")
            out[k] = _encode(v, bufs)
        return out
    if obj is None or isinstance(obj, (bool, int, float, str)):
        return obj
    try:                                           # torch tensors and other array-likes
        return _encode(np.asarray(obj), bufs)
    except Exception as e:                         # pragma: no cover
        raise WireError(f"cannot encode {type(obj).__name__}") from e


def _decode(obj: Any, bufs: List[np.ndarray]):
    if isinstance(obj, list):
        return [_decode(x, bufs) for x in obj]
    if isinstance(obj, dict):
        if "__nd__" in obj and set(obj) == {"__nd__", "dtype", "shape"}:
            return bufs[int(obj["__nd__"])]
        return {k: _decode(v, bufs) for k, v in obj.items()}
    return obj


def _recv_exact(sock: socket.socket, n: int) -> bytes:
    view = bytearray(n)
    got = 0
    while got < n:
        r = sock.recv_into(memoryview(view)[got:], n - got)
        if r == 0:
            raise EOFError("connection closed")
        got += r
    return bytes(view)


def send_msg(sock: socket.socket, obj: Any):
    bufs: List[np.ndarray] = []
    body = _encode(obj, bufs)
    header = json.dumps({"body": body, "sizes": [int(b.nbytes) for b in bufs]}).encode()
    sock.sendall(struct.pack("<I", len(header)) + header)
    for b in bufs:
        if b.nbytes:
            sock.sendall(memoryview(b).cast("B"))


def recv_msg(sock: socket.socket) -> Any:
    (hl,) = struct.unpack("<I", _recv_exact(sock, 4))
    if hl > MAX_HEADER:
        raise WireError("header too large")
    head = json.loads(_recv_exact(sock, hl).decode())
    if not isinstance(head, dict) or "body" not in head or not isinstance(head.get("sizes"), list):
        raise WireError("malformed frame")
    metas: List[Tuple[str, list]] = []

    def collect(o):
        if isinstance(o, list):
            for x in o:
                collect(x)
        elif isinstance(o, dict):
            if "__nd__" in o and set(o) == {"__nd__", "dtype", "shape"}:
                i = int(o["__nd__"])
                while len(metas) <= i:
                    metas.append(None)
                metas[i] = (str(o["dtype"]), [int(x) for x in o["shape"]])
            else:
                for v in o.values():
                    collect(v)

    collect(head["body"])
    if len(metas) != len(head["sizes"]) or any(m is None for m in metas):
        raise WireError("buffer table does not match the header")
    bufs = []
    for (dt, shape), size in zip(metas, head["sizes"]):
        if dt not in _DTYPES or size < 0 or size > MAX_BUFFER:
            raise WireError("bad buffer descriptor")
        want = int(np.prod(shape, dtype=np.int64)) * np.dtype(dt).itemsize if shape else np.dtype(dt).itemsize
        if want != size:
            raise WireError("buffer size does not match dtype/shape")
        raw = _recv_exact(sock, size) if size else b""
        bufs.append(np.frombuffer(raw, dtype=dt).reshape(shape).copy() if size else np.zeros(shape, dtype=dt))
    return _decode(head["body"], bufs)


# ------------------------------------------------------------------------------------------------ authentication
def load_secret(config: Optional[dict] = None) -> Optional[bytes]:
    """Per-deployment secret: ``GW2V_SERVER_SECRET`` (or ``GW2V_SERVER_SECRET_FILE``), else ``secret`` /
    ``secret_file`` of the server config.  None when nothing is configured."""
    s = os.environ.get("GW2V_SERVER_SECRET")
    if s:
        return s.encode()
    sf = os.environ.get("GW2V_SERVER_SECRET_FILE") or (config or {}).get("secret_file")
    if sf:
        with open(sf, "rb") as f:
            return f.read().strip()
    s = (config or {}).get("secret")
    return s.encode() if s else None


def new_secret() -> bytes:
    return secrets.token_hex(32).encode()


def server_handshake(sock: socket.socket, secret: bytes):
    """Challenge / response; raises AuthError (and the caller drops the connection) on a wrong secret."""
    nonce = secrets.token_bytes(32)
    sock.sendall(MAGIC + nonce)
    reply = _recv_exact(sock, 32)
    want = hmac.new(secret, b"client" + nonce, hashlib.sha256).digest()
    if not hmac.compare_digest(reply, want):
        raise AuthError("authentication failed")
    sock.sendall(hmac.new(secret, b"server" + nonce, hashlib.sha256).digest())


def client_handshake(sock: socket.socket, secret: bytes):
    head = _recv_exact(sock, len(MAGIC) + 32)
    if head[:len(MAGIC)] != MAGIC:
        raise WireError("not a glint_word2vec_b200 shard server")
    nonce = head[len(MAGIC):]
    sock.sendall(hmac.new(secret, b"client" + nonce, hashlib.sha256).digest())
    proof = _recv_exact(sock, 32)
    if not hmac.compare_digest(proof, hmac.new(secret, b"server" + nonce, hashlib.sha256).digest()):
        raise AuthError("server failed to authenticate")
