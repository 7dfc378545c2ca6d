"""Tracing / profiling hooks (SURVEY.md 5.1 -- the reference has none).

* ``nvtx_range(name)``: NVTX ranges around the step phases so an Nsight timeline shows
  stage / sub-sample / fused-step / drain; a no-op without CUDA.
* ``StepTimer``: CUDA-event timing of a region on the launching stream (device time, the
  same method bench.py uses), accumulating per-label totals for the metrics JSONL.
* the fused multi-GPU kernels additionally accumulate the *exposed* all-reduce wait in
  device memory (``CudaShardOps.timing``, %globaltimer around the flag spin).
"""
from __future__ import annotations

import contextlib
import time
from collections import defaultdict
from typing import Dict

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class StepTimer:
    """Accumulates device time (CUDA events) or wall time (CPU) per label."""

    def __init__(self, device: torch.device):
        self.cuda = device.type == "cuda"
        self.device = device
        self.totals_ms: Dict[str, float] = defaultdict(float)
        self.counts: Dict[str, int] = defaultdict(int)
        self._pending = []
        self._pool = []             # recycled CUDA events (creating one costs ~0.4 ms)

    @contextlib.contextmanager
    def region(self, label: str):
        if self.cuda:
            e0 = self._pool.pop() if self._pool else torch.cuda.Event(enable_timing=True)
            e1 = self._pool.pop() if self._pool else torch.cuda.Event(enable_timing=True)
            e0.record()
            with nvtx_range(label):
                yield
            e1.record()
            self._pending.append((label, e0, e1))
        else:
            t0 = time.perf_counter()
            yield
            self.totals_ms[label] += (time.perf_counter() - t0) * 1e3
            self.counts[label] += 1

    def poll(self):
        """Harvest the regions whose end event has completed, without synchronising (safe inside a launch loop)."""
        if self.cuda:
            while self._pending and self._pending[0][2].query():
                label, e0, e1 = self._pending.pop(0)
                self.totals_ms[label] += e0.elapsed_time(e1)
                self.counts[label] += 1
                self._pool += [e0, e1]
        return {k: {"ms": v, "n": self.counts[k]} for k, v in self.totals_ms.items()}

    def flush(self):
        if self.cuda and self._pending:
            torch.cuda.synchronize(self.device)
            for label, e0, e1 in self._pending:
                self.totals_ms[label] += e0.elapsed_time(e1)
                self.counts[label] += 1
                self._pool += [e0, e1]
            self._pending.clear()
        return {k: {"ms": v, "n": self.counts[k]} for k, v in self.totals_ms.items()}
