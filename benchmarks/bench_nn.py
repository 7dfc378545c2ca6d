#!/usr/bin/env python
"""Nearest-neighbour lookup benchmark (BASELINE.json config 5: findSynonyms on a 10M x 512 column-sharded syn0).

    python benchmarks/bench_nn.py --gpus N [--vocab 10000000 --dim 512 --queries 64 --k 10]

Device-timed (CUDA events, max over ranks).  Reports queries/s for the tcgen05 (tf32 screen + fp32 re-rank)
path and for the exact CUDA-core path, plus the achieved fraction of the HBM roofline of one sweep over the
shard (V * K * 4 bytes per batch).  The reference does this with one sgemv per query on every server and an
O(V) single-threaded loop + priority queue on the Spark driver (MLLIB:583-630).
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--vocab", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--queries", type=int, default=64)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    if args.gpus > 1 and "RANK" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 1000),
               os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from glint_word2vec_b200.models.engine import ShardEngine
    from glint_word2vec_b200.models.sgns import SGNSConfig
    from glint_word2vec_b200.parallel.comm import Comm, TorchDistComm
    eng = ShardEngine(SGNSConfig(args.vocab, args.dim, seed=3), comm=TorchDistComm() if world > 1 else Comm(), device=dev)
    eng.init_weights()
    eng.norms()
    q = torch.randn(args.queries, args.dim, generator=torch.Generator().manual_seed(1))
    out = {"metric": "findSynonyms queries/sec (device-timed, max over ranks)", "n_gpus": world,
           "config": {"vocab": args.vocab, "dim": args.dim, "queries_per_batch": args.queries, "k": args.k,
                      "cols_per_gpu": eng.shard.cols}}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0) * 1e9
    results = {}
    # select: row-sharded replica + score GEMM with in-epilogue selection (ops/nn.py, csrc/nn_select.cu) -- the product;
    # dense_*: the round-1 path ([Q, V] partial scores reduce-scattered over NVLink, then top-k), kept as fallback
    for name, sel, tcm in (("select_tcgen05", "1", "1"), ("dense_tcgen05_tf32_rerank", "0", "1"), ("dense_exact_fp32", "0", "0")):
        os.environ["GW2V_NN_SELECT"] = sel
        os.environ["GW2V_NN_TC"] = tcm
        for _ in range(3):
            res = eng.top_k(q, args.k)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            eng.top_k(q, args.k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if sel == "1":
            nn = eng._cuda.nn_index()
            sweep_bytes = nn.rows * nn.Kp * 4 * (1.0 + 1.0 / 64)
            extra = {"rows_per_gpu": nn.rows, "row_floats": nn.Kp, "overflows": nn.overflows}
            results[name] = res
        else:
            sweep_bytes = args.vocab * eng.shard.cols * 4
            extra = {}
            results[name] = res
        out[name] = dict({"ms_per_batch": ms, "queries_per_sec": args.queries / (ms * 1e-3),
                          "hbm_fraction_of_measured_copy_bw": sweep_bytes / (ms * 1e-3) / hbm}, **extra)
    a, b = results["select_tcgen05"], results["dense_exact_fp32"]
    out["select_agrees_with_exact"] = {"idx_match": float((a[0] == b[0]).float().mean()),
                                       "max_sim_diff": float((a[1] - b[1]).abs().max())}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
