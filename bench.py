#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): SGNS training throughput in word-pairs/sec
at vocab 10M, dim 512, neg 5, window 5, column-sharded over N B200 GPUs.

    python bench.py --gpus N --steps K --warmup W [--impl fused|baseline|reference]

For N > 1 launch under torchrun (the driver does); a bare ``python bench.py --gpus N``
re-launches itself under ``torch.distributed.run`` on 127.0.0.1.

One JSON line on rank 0.  ``value`` is the whole-job device-timed throughput of
the fused sm_100a step (tokens already on the device); ``e2e`` is the same
metric through the public ``ShardEngine.train_step_async`` API with per-step pinned
host -> device input copies and a device -> host read of the step statistics.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "word-pairs/sec (whole box, device-timed, max over ranks)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fused", choices=["fused", "baseline", "reference"])
    ap.add_argument("--vocab", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--neg", type=int, default=5)
    ap.add_argument("--window", type=int, default=5)
    ap.add_argument("--batch", type=int, default=131072, help="tokens (centre words) per step, global")
    ap.add_argument("--sent-len", type=int, default=1000)
    ap.add_argument("--zipf", type=float, default=1.0)
    ap.add_argument("--subsample", default="word2vec", choices=["reference", "word2vec"],
                    help="word2vec = the intended formula of MLLIB:375-377 at --subsample-ratio; reference = the "
                         "reference's effective behaviour (integer-division bug: nothing is dropped)")
    ap.add_argument("--subsample-ratio", type=float, default=1e-4)
    ap.add_argument("--tile-negatives", type=int, default=64, help="shared negatives per 128-centre tile (neg-sharing tile)")
    ap.add_argument("--neg-sharing", default="pair", choices=["pair", "centre", "tile"],
                    help="pair = n private negatives per pair (reference semantics, the headline); centre = the n "
                         "negatives of a centre are shared by its pairs (optional mode)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--lr", type=float, default=0.025)
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index: int, period: float = 0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_ev = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_ev.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_ev.set()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def main():
    args = parse()
    if args.impl == "reference":
        print(json.dumps({"impl": "reference",
                          "unavailable": "reference is a Scala/sbt Spark+Glint project (no setup.py/pyproject, needs "
                                         "JVM+sbt+network fetch of the Glint fork; contains no GPU code) - pip "
                                         "install of /root/reference fails: not installable"}))
        return 0
    if args.gpus > 1 and "RANK" not in os.environ:
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py needs a GPU"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from glint_word2vec_b200.data.sampler import build_alias, zipf_counts, zipf_tokens
    from glint_word2vec_b200.models.engine import EngineOptions, ShardEngine
    from glint_word2vec_b200.models.sgns import SGNSConfig
    from glint_word2vec_b200.parallel.comm import Comm, TorchDistComm

    comm = TorchDistComm() if world > 1 else Comm()
    cfg = SGNSConfig(args.vocab, args.dim, args.window, args.neg, seed=1, neg_sharing=args.neg_sharing,
                     tile_negatives=args.tile_negatives)
    opts = EngineOptions(subsample_mode=args.subsample, subsample_ratio=args.subsample_ratio)
    eng = ShardEngine(cfg, comm=comm, device=dev, options=opts)
    eng.init_weights()
    B = args.batch
    counts = zipf_counts(args.vocab, 200 * B, args.zipf)
    eng.set_noise(counts)
    stream_alias = build_alias(counts.astype(np.float64))
    W, K = args.warmup, args.steps
    n_steps = W + K
    n_e2e = 0 if args.no_e2e or args.impl != "fused" else (W + K)
    toks = zipf_tokens(stream_alias, (n_steps + n_e2e) * B, seed=2024)
    sid_step = (np.arange(B) // args.sent_len).astype(np.int32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    ops = eng._cuda
    result = {"metric": METRIC, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": W,
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "fp32", "data": "synthetic Zipf(%.2g) token stream, random-init embeddings" % args.zipf,
              "impl": args.impl}

    sampler = ClockSampler(local_rank)
    # ------------------------------------------------------------------ device-timed kernel path
    tok_dev = torch.from_numpy(toks[:n_steps * B]).to(dev)
    sid_dev = torch.from_numpy(sid_step).to(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats_keep = []
    if args.impl == "fused":
        def run_step(s):
            return ops.train_step_device(tok_dev[s * B:(s + 1) * B], sid_dev, B, s * B, 0, args.lr)
    else:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        from nccl_sgns import BaselineShard
        bl = BaselineShard(eng)
        pre = [bl.enumerate(toks[s * B:(s + 1) * B], sid_step, s * B, 0) for s in range(n_steps)]

        def run_step(s):
            n = bl.step(None, None, s * B, 0, args.lr, pairs_host=pre[s])
            return torch.tensor([float(n), 0, 0, B])
    for s in range(W):
        run_step(s)
    barrier()
    wait0 = int(ops.timing[0].item()) if (world > 1 and ops.timing is not None) else 0
    launches0 = ops.launches
    sampler.start()
    t_host0 = time.perf_counter()
    ev0.record()
    for s in range(W, n_steps):
        stats_keep.append(run_step(s).clone())
    ev1.record()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / K          # host time to QUEUE one step (diagnostic)
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    if world > 1 and ops.timing is not None and ops._xchg is not None:
        # in-kernel %globaltimer time spent polling for the peers' partial dots, timed region only:
        # mean over the exchanging warps, per step (a mini-batch of the reference = 50 centres of this step)
        warps = ops._xchg["grid"] * 8
        wait_us = (int(ops.timing[0].item()) - wait0) / 1e3 / max(1, warps) / K
        result["exposed_allreduce_us_per_step"] = max_over_ranks(wait_us)
        result["exposed_allreduce_fraction_of_step"] = result["exposed_allreduce_us_per_step"] / (ms / K * 1e3)
    launches = ops.launches - launches0
    pairs = float(sum(float(x[0]) for x in stats_keep))
    value = pairs / (ms * 1e-3)
    result.update({"value": value, "ms_per_step": ms / K, "gpu_launches": launches if args.impl == "fused" else 0,
                   "pairs_per_step": pairs / K, "host_enqueue_ms_per_step": host_enqueue_ms})

    # ------------------------------------------------------------------ end to end through the public API
    if n_e2e:
        pin_tok = [torch.from_numpy(toks[(n_steps + s) * B:(n_steps + s + 1) * B].copy()).pin_memory()
                   for s in range(n_e2e)]
        pin_sid = torch.from_numpy(sid_step.copy()).pin_memory()
        # software-pipelined by one step, exactly like trainer.train(): queue step s (its pinned-host -> device copy
        # runs on the copy stream), then read step s-1's statistics (asynchronous D2H into a pinned ring + event)
        for s in range(W):
            eng.train_step_async(pin_tok[s], pin_sid, (n_steps + s) * B, 0, args.lr).result()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2e_pairs = 0.0
        prev = None
        e0.record()
        for s in range(W, n_e2e):
            h = eng.train_step_async(pin_tok[s], pin_sid, (n_steps + s) * B, 0, args.lr)   # H2D: tokens + sentence ids
            if prev is not None:
                e2e_pairs += float(prev.result()[0])                                     # D2H: pairs, loss, ...
            prev = h
        e2e_pairs += float(prev.result()[0])
        e1.record()
        barrier()
        ms2 = max_over_ranks(e0.elapsed_time(e1))
        result["e2e"] = {"value": e2e_pairs / (ms2 * 1e-3), "unit": "pairs/s", "ms_per_step": ms2 / K,
                         "h2d_bytes_per_step": int(B * 4 * 2), "d2h_bytes_per_step": 16,
                         "api": "ShardEngine.train_step_async(pinned tokens, pinned sent_id) -> handle.result() "
                                "(every step: H2D of its inputs, D2H of its statistics; read-back lags one step)"}
    sampler.stop()
    sampler.join(timeout=2)
    result["clocks"] = sampler.summary()
    if world > 1 and ops.timing is not None:
        result["exposed_allreduce_wait_ns_total"] = int(ops.timing[0].item())
    result["config"] = {
        "model": "SGNS word2vec", "vocab": args.vocab, "dim": args.dim, "neg": args.neg, "window": args.window,
        "global_batch": B, "seq_len": args.sent_len, "parallelism": f"column-shard x{world}",
        "cols_per_gpu": eng.shard.cols, "window_mode": cfg.window_mode,
        "subsample": args.subsample + (" t=%g" % args.subsample_ratio if args.subsample == "word2vec" else " (inert)"),
        "pairs_counted": "trained (centre, context) pairs after sub-sampling",
        "neg_sharing": "pair (n private negatives per (centre, context) pair)" if args.neg_sharing == "pair"
        else ("centre (the n negatives of a centre are shared by its pairs) - NOT the reference semantics"
              if args.neg_sharing == "centre" else
              "tile (%d negatives shared by each tile of 128 centres, weighted m_i*n/N; tcgen05 kernel) - NOT the "
              "reference semantics" % args.tile_negatives),
        "l2": "inputs (2 x %.1f GB embedding shards per GPU) far larger than the 126 MB L2; no flush needed"
              % (args.vocab * eng.shard.cols * 4 / 1e9),
    }
    if rank == 0:
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
