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


class Checkpointer:
    def __init__(self, engine: ShardEngine, directory: str, every_steps: int, hyper: dict, keep: int = 2):
        self.engine = engine
        self.dir = directory
        self.every = max(1, int(every_steps))
        self.hyper = dict(hyper)
        self.keep = keep
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
                         world=eng.comm.world)
            with open(os.path.join(path, "state.json"), "w") as f:
                json.dump(state, f)
            tmp = os.path.join(self.dir, "LATEST.tmp")
            with open(tmp, "w") as f:
                f.write(name)
            os.replace(tmp, os.path.join(self.dir, "LATEST"))
            olds = sorted(d for d in os.listdir(self.dir) if d.startswith("ckpt-"))
            for d in olds[:-self.keep]:
                shutil.rmtree(os.path.join(self.dir, d), ignore_errors=True)
        eng.comm.barrier()


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
    ck = Checkpointer(eng, directory, every_steps, {k: st[k] for k in ("learning_rate", "num_iterations",
                                                                       "train_words", "step_tokens")}) \
        if every_steps > 0 else None
    rep = trainer.train(eng, corpus, st["learning_rate"], st["num_iterations"], st["train_words"],
                        metrics_path=metrics_path, checkpoint_fn=ck,
                        start_iteration=st["iteration"], start_step=st["next_step"])
    return eng, rep


def train_with_checkpoints(engine: ShardEngine, corpus: EncodedCorpus, learning_rate: float, num_iterations: int,
                           train_words: int, directory: str, every_steps: int, metrics_path: Optional[str] = None):
    step_tokens = trainer.auto_step_tokens(engine, corpus.num_tokens)
    engine.opts.step_tokens = step_tokens                      # freeze the partition for a later resume
    ck = Checkpointer(engine, directory, every_steps,
                      dict(learning_rate=learning_rate, num_iterations=num_iterations, train_words=train_words,
                           step_tokens=step_tokens))
    return trainer.train(engine, corpus, learning_rate, num_iterations, train_words, metrics_path=metrics_path,
                         checkpoint_fn=ck)
