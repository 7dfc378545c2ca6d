"""On-disk layout of the column-sharded matrices and training checkpoints.

Reference: ``matrix.save(path, hadoopConf)`` makes every parameter server write
its slice plus Glint metadata under ``<path>`` (MLLIB:494), and
``client.loadWord2vecMatrix(path, ...)`` reads it back with however many
servers the loading cluster has (MLLIB:716-722, quirk Q12).  Glint's own blob
layout is not knowable offline, so this is our documented one::

    <path>/matrix/meta.json               format, V, d, shard table, SGNS config
    <path>/matrix/syn0.<r>of<S>.npy       [V, real_cols_r] float32, columns col_start_r..
    <path>/matrix/syn1.<r>of<S>.npy       same for syn1neg (optional -> retrainable)
    <path>/matrix/_SUCCESS

Loading never assumes the saved shard count equals the loading world size:
each rank memory-maps every saved slice and keeps the overlap with its own
column range (re-sharding, Q12).
"""
from __future__ import annotations

import json
import os
from typing import Optional

import numpy as np

from .engine import EngineOptions, ShardEngine
from .sgns import SGNSConfig

FORMAT = "glint_word2vec_b200.matrix"
VERSION = 1


def _matrix_dir(path: str) -> str:
    return os.path.join(path, "matrix")


CHUNK_BYTES = 256 << 20          # rows move between device and file in pieces of this size


def _save_npy_streamed(t, filename: str):
    """Write a 2-D float32 tensor (any device, possibly a column slice) as a standard ``.npy`` file without ever
    holding more than two chunks on the host: device -> pinned staging buffer -> file, the copy of chunk i+1
    overlapping the write of chunk i."""
    import torch
    from numpy.lib import format as npformat
    rows, cols = int(t.shape[0]), int(t.shape[1])
    with open(filename, "wb") as f:
        npformat.write_array_header_1_0(f, {"descr": "<f4", "fortran_order": False, "shape": (rows, cols)})
        if rows == 0 or cols == 0:
            return
        step = max(1, CHUNK_BYTES // (cols * 4))
        cuda = t.is_cuda
        bufs = [torch.empty(min(step, rows), cols, dtype=torch.float32, pin_memory=cuda) for _ in range(2)]
        evs = [torch.cuda.Event() if cuda else None for _ in range(2)]

        def fetch(i, lo):
            n = min(step, rows - lo)
            bufs[i][:n].copy_(t[lo:lo + n], non_blocking=cuda)
            if cuda:
                evs[i].record()
            return n
        lo, i = 0, 0
        n = fetch(i, lo)
        while lo < rows:
            nxt = lo + n
            n_next = fetch(1 - i, nxt) if nxt < rows else 0
            if cuda:
                evs[i].synchronize()
            f.write(memoryview(bufs[i][:n].numpy()))          # contiguous row block of the staging buffer
            lo, n, i = nxt, n_next, 1 - i


def save_matrix(engine: ShardEngine, path: str, extra: Optional[dict] = None):
    """Collective: every rank writes its slices, rank 0 writes the metadata."""
    mdir = _matrix_dir(path)
    comm = engine.comm
    if comm.rank == 0:
        os.makedirs(mdir, exist_ok=True)
    comm.barrier()
    sh = engine.shard
    entry = {"rank": sh.rank, "col_start": sh.col_start, "cols": sh.real_cols}
    for name, t in engine.shard_tensors().items():
        fn = f"{name}.{sh.rank:02d}of{sh.world:02d}.npy"
        _save_npy_streamed(t, os.path.join(mdir, fn))
        entry[name] = fn
    entries = comm.gather_objects(entry, dst=0)
    if comm.rank == 0:
        meta = {
            "format": FORMAT, "version": VERSION,
            "vocab_size": engine.cfg.vocab_size, "vector_size": engine.cfg.vector_size,
            "num_shards": sh.world, "dtype": "float32",
            "shards": sorted(entries, key=lambda e: e["rank"]),
            "config": engine.cfg.to_dict(),
        }
        if extra:
            meta["extra"] = extra
        with open(os.path.join(mdir, "meta.json"), "w") as f:
            json.dump(meta, f, indent=1)
        with open(os.path.join(mdir, "_SUCCESS"), "w"):
            pass
    comm.barrier()


def read_meta(path: str) -> dict:
    with open(os.path.join(_matrix_dir(path), "meta.json")) as f:
        meta = json.load(f)
    if meta.get("format") != FORMAT:
        raise ValueError(f"{path}: not a {FORMAT} directory")
    return meta


def load_matrix(path: str, comm, device=None, options: Optional[EngineOptions] = None,
                with_syn1: bool = True) -> ShardEngine:
    """Collective: build an engine for the loading world size and fill it from
    whatever shard layout was saved."""
    meta = read_meta(path)
    cfg = SGNSConfig(**meta["config"])
    engine = ShardEngine(cfg, comm=comm, device=device, options=options)
    mdir = _matrix_dir(path)
    sh = engine.shard
    lo, hi = sh.col_start, sh.col_start + sh.real_cols
    names = ["syn0"] + (["syn1"] if with_syn1 else [])
    for name in names:
        for e in meta["shards"]:
            if name not in e:
                continue
            s_lo, s_hi = e["col_start"], e["col_start"] + e["cols"]
            if e["cols"] == 0 or s_hi <= lo or s_lo >= hi:
                continue
            block = np.load(os.path.join(mdir, e[name]), mmap_mode="r")
            engine.load_columns(name, s_lo, block)
    # make sure both matrices exist even for empty overlaps / missing syn1
    import torch
    for attr in ("syn0", "syn1"):
        if getattr(engine, attr) is None:
            setattr(engine, attr, torch.zeros(cfg.vocab_size, sh.cols, dtype=torch.float32,
                                               device=engine.device))
    return engine
