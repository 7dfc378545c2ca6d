"""Mid-training checkpoints and resume.

The reference can only save a finished model (MLLIB:493-498); a lost parameter
server loses its column slice and the whole job.  Here the
training state is just two matrices plus a handful of counters, so a periodic
checkpoint is cheap and is the recovery story for a failed rank:

    <dir>/ckpt-<k>-<step>/matrix/...     column shards (models/matrix_io.py, syn0 + syn1neg)
    <dir>/ckpt-<k>-<step>/state.json     iteration, step, step_tokens, hyper-parameters, counters
    <dir>/LATEST                         name of the newest complete checkpoint (atomic rename)

Because every random decision is a pure function of (seed, iteration, position),
resuming at (iteration, step) reproduces exactly the windows / negatives /
sub-sampling the uninterrupted run would have used.  Resume works with a
different number of shards than the run that wrote the checkpoint (re-sharding).
"""
from __future__ import annotations

import json
import os
import shutil
import time
from typing import Optional

from ..data.corpus import EncodedCorpus
from . import matrix_io, trainer
from .engine import EngineOptions, ShardEngine


def fingerprint(engine: ShardEngine, corpus: EncodedCorpus) -> dict:
    """What a checkpoint belongs to: the model configuration and a cheap digest of the corpus.  ``resume`` refuses a
    checkpoint whose fingerprint differs (a stale directory of another run, another corpus, another seed...)."""
    import hashlib
    import numpy as np
    h = hashlib.sha256()
    n = corpus.num_tokens
    for lo in (0, max(0, n // 2 - (1 << 16)), max(0, n - (1 << 17))):
        h.update(np.ascontiguousarray(corpus.tokens[lo:lo + (1 << 17)], dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(corpus.offsets[:1 << 16], dtype=np.int64).tobytes())
    cfg = engine.cfg
    return {"tokens": int(n), "sentences": int(corpus.num_sentences), "digest": h.hexdigest()[:32],
            "vocab_size": int(cfg.vocab_size), "vector_size": int(cfg.vector_size), "window": int(cfg.window),
            "negatives": int(cfg.negatives), "seed": int(cfg.seed), "window_mode": cfg.window_mode,
            "neg_sharing": cfg.neg_sharing, "subsample_mode": engine.opts.subsample_mode,
            "subsample_ratio": float(engine.opts.subsample_ratio)}


class Checkpointer:
    def __init__(self, engine: ShardEngine, directory: str, every_steps: int, hyper: dict, keep: int = 2,
                 fp: Optional[dict] = None, run_id: Optional[str] = None):
        import uuid
        self.engine = engine
        self.dir = directory
        self.every = max(1, int(every_steps))
        self.hyper = dict(hyper)
        self.keep = max(1, keep)
        self.fp = fp
        self.run_id = run_id or uuid.uuid4().hex[:16]
        self._mine = []                      # checkpoints written by THIS object, oldest first: the only ones it prunes
        self._count = 0
        if engine.comm.rank == 0:
            os.makedirs(directory, exist_ok=True)

    def __call__(self, iteration: int, next_step: int):
        self._count += 1
        if self._count % self.every:
            return
        self.save(iteration, next_step)

    def save(self, iteration: int, next_step: int):
        eng = self.engine
        name = f"ckpt-{iteration:04d}-{next_step:08d}"
        path = os.path.join(self.dir, name)
        matrix_io.save_matrix(eng, path)                       # collective
        if eng.comm.rank == 0:
            state = dict(self.hyper, iteration=iteration, next_step=next_step, time=time.time(),
                         world=eng.comm.world, run_id=self.run_id, fingerprint=self.fp)
            with open(os.path.join(path, "state.json"), "w") as f:
                json.dump(state, f)
            tmp = os.path.join(self.dir, "LATEST.tmp")
            with open(tmp, "w") as f:
                f.write(name)
            os.replace(tmp, os.path.join(self.dir, "LATEST"))
            if name in self._mine:
                self._mine.remove(name)
            self._mine.append(name)
            # prune only what this run wrote, oldest first, and never the checkpoint LATEST points at: directories of
            # other runs (whatever their iteration / step numbers) are left alone
            while len(self._mine) > self.keep:
                old = self._mine.pop(0)
                if old != name:
                    shutil.rmtree(os.path.join(self.dir, old), ignore_errors=True)
        eng.comm.barrier()


def prepare_directory(engine: ShardEngine, directory: str, overwrite: bool = False):
    """A fresh (non-resuming) run must not mix with the checkpoints of an earlier one: refuse a directory that
    already holds a LATEST pointer unless ``overwrite`` (then the old checkpoints are removed)."""
    if engine.comm.rank == 0 and os.path.isdir(directory):
        stale = [d for d in os.listdir(directory) if d.startswith("ckpt-") or d == "LATEST"]
        if stale and not overwrite:
            raise FileExistsError(f"{directory} already holds checkpoints of another run ({len(stale)} entries): resume "
                                  "it (resume=true), point checkpoint_dir elsewhere, or pass checkpoint_overwrite=true")
        for d in stale:
            p = os.path.join(directory, d)
            shutil.rmtree(p, ignore_errors=True) if os.path.isdir(p) else os.remove(p)
    engine.comm.barrier()


def check_fingerprint(state: dict, fp: dict):
    old = state.get("fingerprint")
    if old is None:
        return                                                   # checkpoint of an older format: nothing to compare
    diff = {k: (old.get(k), fp.get(k)) for k in fp if old.get(k) != fp.get(k)}
    if diff:
        raise ValueError("checkpoint does not belong to this run (corpus / configuration differ): "
                         + ", ".join(f"{k}: saved {a!r} != current {b!r}" for k, (a, b) in diff.items()))


def latest(directory: str) -> Optional[str]:
    p = os.path.join(directory, "LATEST")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        name = f.read().strip()
    path = os.path.join(directory, name)
    return path if os.path.exists(os.path.join(path, "state.json")) else None


def resume(directory: str, corpus: EncodedCorpus, counts, comm, device=None,
           options: Optional[EngineOptions] = None, every_steps: int = 0, metrics_path: Optional[str] = None):
    """Collective: load the newest checkpoint (re-sharding if needed) and finish the run.

    Returns ``(engine, TrainReport)``."""
    path = latest(directory)
    if path is None:
        raise FileNotFoundError(f"no complete checkpoint under {directory}")
    with open(os.path.join(path, "state.json")) as f:
        st = json.load(f)
    opts = options if options is not None else EngineOptions()
    opts.step_tokens = int(st["step_tokens"])                  # the step partition must be the same
    eng = matrix_io.load_matrix(path, comm, device, opts)
    eng.set_noise(counts)
    fp = fingerprint(eng, corpus)
    check_fingerprint(st, fp)
    ck = Checkpointer(eng, directory, every_steps, {k: st[k] for k in ("learning_rate", "num_iterations",
                                                                       "train_words", "step_tokens")},
                      fp=fp, run_id=st.get("run_id")) if every_steps > 0 else None
    rep = trainer.train(eng, corpus, st["learning_rate"], st["num_iterations"], st["train_words"],
                        metrics_path=metrics_path, checkpoint_fn=ck,
                        start_iteration=st["iteration"], start_step=st["next_step"])
    return eng, rep


def train_with_checkpoints(engine: ShardEngine, corpus: EncodedCorpus, learning_rate: float, num_iterations: int,
                           train_words: int, directory: str, every_steps: int, metrics_path: Optional[str] = None,
                           overwrite: bool = False):
    prepare_directory(engine, directory, overwrite)
    step_tokens = trainer.auto_step_tokens(engine, corpus.num_tokens)
    engine.opts.step_tokens = step_tokens                      # freeze the partition for a later resume
    ck = Checkpointer(engine, directory, every_steps,
                      dict(learning_rate=learning_rate, num_iterations=num_iterations, train_words=train_words,
                           step_tokens=step_tokens), fp=fingerprint(engine, corpus))
    return trainer.train(engine, corpus, learning_rate, num_iterations, train_words, metrics_path=metrics_path,
                         checkpoint_fn=ck)
