"""Training driver: the client side of the reference's hot loop (C7/C8).

Reference: ``doFit`` iterates ``numIterations`` times over the cached
sentences, decays alpha every >10 000 words and issues one ``dotprod`` +
``adjust`` round trip per <=50-centre mini-batch (MLLIB:364-433).  Here every
rank runs this same loop in lock-step (SPMD): the token stream is replicated
(or regenerated) on every rank, one ``ShardEngine.train_step`` covers
thousands of mini-batches, and alpha follows the closed form with true global
progress.
"""
from __future__ import annotations

import json
import logging
import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np
import torch

from ..data.corpus import EncodedCorpus, iter_steps
from ..utils.tracing import StepTimer
from . import sgns
from .engine import ShardEngine

log = logging.getLogger("glint_word2vec_b200.trainer")


@dataclass
class TrainReport:
    iterations: int = 0
    steps: int = 0
    words: int = 0
    pairs: int = 0
    loss_per_pair: float = float("nan")
    max_abs_dot: float = 0.0
    seconds: float = 0.0
    final_alpha: float = 0.0
    history: List[dict] = field(default_factory=list)
    device_ms: dict = field(default_factory=dict)     # per-phase device time {"sgns_step": {"ms": ..., "n": steps}}

    @property
    def pairs_per_sec(self) -> float:
        return self.pairs / self.seconds if self.seconds > 0 else float("nan")


def auto_step_tokens(engine: ShardEngine, corpus_tokens: int) -> int:
    """Tokens per device step.

    On a GPU every centre of a step is in flight at once (thousands of warps / 148 tiles), so a row that occurs c
    times in the step receives c summed updates computed from nearly the same stale values -- the asynchronous-SGD
    hazard the reference warns about (README.md:17-19: "sensitive to very frequent words ... exploding gradients").

    * hot-row damping on (default, ``hot_row_cap > 0``): the hot rows are protected whatever the step size, so the step
      is sized for statistical efficiency -- at least ~512 sequential steps per pass over a small corpus -- and capped
      at 131 072 tokens (the benchmark step) for large ones;
    * damping off: the hottest word (after sub-sampling) is expected at most ``max_hot_updates`` times per step.
    Clamped to [256, 256k]; ``step_tokens`` overrides both."""
    opts = engine.opts
    if opts.step_tokens > 0:
        return opts.step_tokens
    if not engine.is_cuda:
        return max(opts.batch_size, min(1 << 16, corpus_tokens))
    if opts.hot_row_cap > 0:
        return int(max(256, min(1 << 17, corpus_tokens // 512)))
    f_max = 1.0
    if engine.alias is not None and engine.keep_thresh is not None and engine.noise_counts is not None:
        eff = engine.noise_counts.astype(np.float64) * (engine.keep_thresh.astype(np.float64) + 1.0) / 2.0 ** 32
        f_max = float(eff.max() / max(eff.sum(), 1.0))
    step = int(opts.max_hot_updates / max(f_max, 1e-9))
    return int(max(256, min(1 << 18, step)))


def train(engine: ShardEngine, corpus: EncodedCorpus, learning_rate: float, num_iterations: int,
          train_words: Optional[int] = None, log_every_words: int = 10000,
          metrics_path: Optional[str] = None, checkpoint_fn: Optional[Callable[[int, int], None]] = None,
          start_iteration: int = 0, start_step: int = 0) -> TrainReport:
    """Run ``num_iterations`` passes; every rank calls this with the same corpus."""
    rep = TrainReport()
    n_tok = corpus.num_tokens
    train_words = n_tok if train_words is None else train_words
    total_words = num_iterations * train_words
    step_tokens = auto_step_tokens(engine, n_tok)
    pool = None
    if engine.is_cuda and not engine.unfused:
        # one-time work outside the clock: buffers, damping tables, exchange rings, the pinned producer pool, and a
        # zero-token step that makes the driver load every kernel of the step (lazy module loading: ~0.2 s)
        engine._cuda.prepare(min(step_tokens, max(1, n_tok)))
        pool = _PinnedPool(step_tokens)
        if engine.comm.world == 1:                  # (column shards: the first real step does it; not validated on > 1 GPU)
            engine._cuda.warmup()
        torch.cuda.synchronize(engine.device)
    t0 = time.time()
    pending = []
    last_log = 0
    mf = open(metrics_path, "a") if (metrics_path and engine.comm.rank == 0) else None
    alpha = learning_rate
    # statistics are read back LAG steps late: resolving the newest handle would make the host wait for the step it
    # has just queued (one host<->device round trip per step; measured 6x slower than the engine-level loop)
    lag = 2 if engine.is_cuda else 0
    timer = StepTimer(engine.device)
    for k in range(start_iteration, num_iterations):
        words_prev = k * train_words
        words_it = 0
        skip = start_step if k == start_iteration else 0
        # steps are cut (memory-mapped token slices, sentence ids) by a background thread, PREFETCH steps ahead
        for si, batch in enumerate(_prefetch(_pinned(iter_steps(corpus, step_tokens), pool), PREFETCH)):
            if si < skip:
                words_it += batch.n_words
                continue
            alpha = sgns.learning_rate(learning_rate, words_prev + words_it, total_words)
            with timer.region("sgns_step"):          # device time of the step's kernels (CUDA events) + NVTX range
                stats = engine.train_step_async(batch.tokens, batch.sent_id, batch.raw_pos0, k, alpha)
            pending.append(stats)
            words_it += batch.n_words
            rep.steps += 1
            # GPU: a metrics record every >= 8 steps (a record per 131 072-token step costs more host time than the step)
            if (words_prev + words_it - last_log > max(log_every_words, lag * 4 * step_tokens) and len(pending) > lag) \
                    or len(pending) >= 64:
                last_log = words_prev + words_it
                ready = pending[:len(pending) - lag]
                del pending[:len(pending) - lag]
                _drain(ready, rep, alpha, words_prev + words_it, mf, engine, t0, timer.poll())
            if checkpoint_fn is not None:
                checkpoint_fn(k, si + 1)
        rep.iterations += 1
        rep.words += words_it
    _drain(pending, rep, alpha, rep.words, mf, engine, t0, timer.flush())
    if engine.is_cuda:
        torch.cuda.synchronize(engine.device)
    rep.seconds = time.time() - t0
    rep.device_ms = timer.flush()
    rep.final_alpha = alpha
    if mf:
        mf.close()
    return rep


PREFETCH = 4


class _PinnedPool:
    """Ring of pinned host buffers shared by all passes of one ``train`` call (allocated before the clock starts:
    32 ``cudaHostAlloc`` calls cost ~0.15 s).  PREFETCH + 2 x staging slots + 4 buffers: the host can be at most that many
    steps ahead of the copy engine (``stage_tokens`` blocks on the slot's free event; the launch loop learns about
    completions one step late), so a buffer is never rewritten before its H2D ran."""

    def __init__(self, step_tokens: int):
        from ..ops.cuda import N_STAGE
        self.n = PREFETCH + 2 * N_STAGE + 4
        self.cap = max(1, int(step_tokens))
        self.bufs = [(torch.empty(self.cap, dtype=torch.int32).pin_memory(), torch.empty(self.cap, dtype=torch.int32).pin_memory())
                     for _ in range(self.n)]
        self.views = [(a.numpy(), b.numpy()) for a, b in self.bufs]      # numpy views: plain memcpy, read-only maps are fine
        self.i = 0


def _pinned(steps, pool: Optional["_PinnedPool"]):
    """GPU engines: copy every step's arrays into pinned host buffers INSIDE the producer thread, so the launch loop
    hands pinned tensors to ``stage_tokens`` (asynchronous H2D straight from them, no host memcpy on the main thread)."""
    if pool is None:
        yield from steps
        return
    from ..data.corpus import StepBatch
    for b in steps:
        t = int(b.tokens.shape[0])
        if t > pool.cap:                              # cannot happen with iter_steps; keep the slow path correct
            yield b
            continue
        pt, ps = pool.bufs[pool.i]
        vt, vs = pool.views[pool.i]
        pool.i = (pool.i + 1) % pool.n
        vt[:t] = b.tokens
        vs[:t] = b.sent_id
        yield StepBatch(pt[:t], ps[:t], b.raw_pos0, b.n_words)


def _prefetch(it, depth: int):
    """Run iterator ``it`` in a daemon thread, ``depth`` items ahead (numpy slicing, memmap page faults and
    ``np.repeat`` release the GIL, so the producer overlaps the launch loop).  Exceptions re-raise in the consumer;
    an abandoned consumer stops the producer at its next hand-over."""
    import queue
    import threading
    q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
    done = object()
    stop = threading.Event()

    def put(x):
        while not stop.is_set():
            try:
                q.put(x, timeout=0.2)
                return True
            except queue.Full:
                continue
        return False

    def run():
        try:
            for x in it:
                if not put(x):
                    return
            put(done)
        except BaseException as exc:          # noqa: BLE001 - handed to the consumer
            put(exc)

    th = threading.Thread(target=run, name="gw2v-steps", daemon=True)
    th.start()
    try:
        while True:
            x = q.get()
            if x is done:
                return
            if isinstance(x, BaseException):
                raise x
            yield x
    finally:
        stop.set()


def _exposed_wait_ns(engine) -> Optional[int]:
    ops = getattr(engine, "_cuda", None) if engine is not None else None
    timing = getattr(ops, "timing", None)
    if timing is None:
        return None
    return int(timing[0].item())


def _drain(pending, rep: TrainReport, alpha, words, mf, engine=None, t0=None, phases=None):
    if not pending:
        return
    vals = [p.result() if hasattr(p, "result") else p for p in pending]      # async step handles
    st = torch.stack([v.to("cpu", torch.float64) if isinstance(v, torch.Tensor) else torch.tensor(v, dtype=torch.float64)
                      for v in vals])
    pending.clear()
    pairs = int(st[:, 0].sum())
    loss = float(st[:, 1].sum())
    maxdot = float(st[:, 2].max())
    tot_pairs = rep.pairs + pairs
    if tot_pairs > 0:
        prev = 0.0 if rep.pairs == 0 or rep.loss_per_pair != rep.loss_per_pair else rep.loss_per_pair * rep.pairs
        rep.loss_per_pair = (prev + loss) / tot_pairs
    rep.pairs = tot_pairs
    rep.max_abs_dot = max(rep.max_abs_dot, maxdot)
    rec = {"words": int(words), "alpha": float(alpha), "pairs": pairs,
           "loss_per_pair": (loss / pairs) if pairs else None, "max_abs_dot": maxdot}
    if t0 is not None:                                   # throughput next to the loss probe
        rec["elapsed_s"] = time.time() - t0
        rec["pairs_per_sec"] = rep.pairs / rec["elapsed_s"] if rec["elapsed_s"] > 0 else None
    if phases:                                           # cumulative device time per phase (utils/tracing.py::StepTimer)
        rec["device_ms"] = {k: round(v["ms"], 3) for k, v in phases.items()}
        rec["device_steps"] = {k: v["n"] for k, v in phases.items()}
    wait = _exposed_wait_ns(engine)
    if wait is not None:                                 # in-kernel time spent polling for the peers' partial dots
        rec["exposed_allreduce_wait_ns_total"] = wait
    rep.history.append(rec)
    # same probe the reference logs every 10k words: wordCount, alpha, a dot product (MLLIB:411-412)
    log.info("wordCount = %d, alpha = %.6g, loss/pair = %s, max|f| = %.4g", words, alpha,
             rec["loss_per_pair"], maxdot)
    if mf:
        mf.write(json.dumps(rec) + "\n")
        mf.flush()
