"""Multi-GPU: the in-kernel NVLink exchange of the fused SGNS step.

World size 2/4/8 must reproduce the single-shard result (same Philox pairs and
negatives; only fp32 summation order and Hogwild timing differ), and the
serving collectives must match.  Skipped on boxes with fewer GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, d, out_dir, n=5):
    import torch.distributed as dist
    from glint_word2vec_b200.data.sampler import zipf_counts
    from glint_word2vec_b200.models import sgns
    from glint_word2vec_b200.models.engine import EngineOptions, ShardEngine
    from glint_word2vec_b200.models.sgns import SGNSConfig
    from glint_word2vec_b200.parallel.comm import TorchDistComm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        v, t = 100000, 6000
        cfg = SGNSConfig(v, d, 5, n, seed=11)
        eng = ShardEngine(cfg, comm=TorchDistComm(), device=dev, options=EngineOptions(subsample_mode="reference", hot_row_cap=0))
        eng.init_weights()
        counts = zipf_counts(v, 10 ** 7, 0.6)
        eng.set_noise(counts)
        # non-trivial output vectors so dots/gradients are not all identical
        full1 = (torch.rand(v, eng.shard.padded_vector_size, generator=torch.Generator().manual_seed(5)) - 0.5) * 0.5
        full1[:, d:] = 0
        eng.syn1 = full1[:, rank * eng.shard.cols:(rank + 1) * eng.shard.cols].contiguous().to(dev)
        eng.syn0 = (eng.syn0 * 20.0).contiguous()
        rng = np.random.default_rng(3)
        steps = []
        for s in range(3):
            tokens = rng.choice(v, size=t, replace=False).astype(np.int32)
            sid = (np.arange(t) // 41).astype(np.int32)
            steps.append((tokens, sid))
        start0 = eng.pull(torch.arange(v)).cpu()
        stats = []
        for s, (tokens, sid) in enumerate(steps):
            stats.append(eng.train_step(tokens, sid, s * t, 0, 0.05).cpu())
        torch.cuda.synchronize(dev)
        got0 = eng.pull(torch.arange(v)).cpu()
        nrm = eng.norms().cpu()
        idx, sim = eng.top_k(got0[:4], 5)
        # serving ops with the collective fused into the kernels (ops/serving.py)
        avg = eng.pull_average(torch.tensor([1, 2, 3, 7, 9]), torch.tensor([0, 3, 3, 5])).cpu()
        mul = eng.multiply(got0[5]).cpu()
        idx16, sim16 = eng.top_k(got0[100:116], 7)          # Q = 16: tcgen05 screening + exact re-rank when K % 32 == 0
        if rank == 0:
            # oracle: dense single-process mini-batch semantics, one mini-batch per step
            ref0, _ = sgns.init_embeddings(v, d, 11)
            ref0 = ref0 * 20.0
            ref1 = full1[:, :d].clone()
            assert torch.allclose(start0, ref0, atol=1e-7)
            pairs = []
            for s, (tokens, sid) in enumerate(steps):
                st = sgns.sgns_minibatch_reference(ref0, ref1, cfg, eng.alias, tokens, sid, s * t, 0, 0.05)
                pairs.append(st.pairs)
            torch.save({"got0": got0, "ref0": ref0, "start0": start0, "pairs": pairs,
                        "stats": torch.stack(stats), "nrm": nrm, "idx": idx, "sim": sim,
                        "avg": avg, "mul": mul, "idx16": idx16, "sim16": sim16,
                        "xchg": eng._cuda.xchg_selftest, "debug": eng._cuda.debug},
                       os.path.join(out_dir, "result.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,d,nvls,serve_fused,debug", [(2, 128, 0, 1, 0), (2, 100, 0, 1, 0), (2, 128, 1, 0, 0),
                                                            (2, 64, 0, 1, 16), (4, 256, 0, 1, 0), (8, 512, 0, 1, 0),
                                                            (8, 512, 1, 1, 16)])
def test_fused_multi_matches_oracle(world, d, nvls, serve_fused, debug, tmp_path, monkeypatch):
    """nvls=1: the batch push uses multimem.st on the NVLS multicast mapping when the driver grants one
    (falls back to per-peer stores otherwise).  serve_fused=0: serving ops as kernel + NCCL collective
    instead of the in-kernel peer pushes.  debug=16: ranks are de-synchronised by random in-kernel delays."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    monkeypatch.setenv("GW2V_NVLS", str(nvls))
    monkeypatch.setenv("GW2V_SERVE_FUSED", str(serve_fused))
    # debug=16: random per-warp delays before every push (flag/slot protocol stress test)
    monkeypatch.setenv("GW2V_DEBUG", str(debug))
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), d, str(tmp_path)), nprocs=world, join=True)
    r = torch.load(os.path.join(tmp_path, "result.pt"))
    assert [int(x) for x in r["stats"][:, 0]] == r["pairs"]
    upd_ref = r["ref0"] - r["start0"]
    upd_got = r["got0"] - r["start0"]
    assert upd_ref.norm() > 0
    assert (upd_got - upd_ref).norm() / upd_ref.norm() < 3e-2
    assert torch.allclose(r["nrm"], r["got0"].norm(dim=1), rtol=1e-4, atol=1e-6)
    # each of the first rows is its own nearest neighbour with cosine 1
    assert r["idx"][:, 0].tolist() == [0, 1, 2, 3]
    assert torch.allclose(r["sim"][:, 0], torch.ones(4), atol=1e-4)
    g = r["got0"]
    want_avg = torch.stack([g[[1, 2, 3]].mean(0), torch.zeros(d), g[[7, 9]].mean(0)])
    assert torch.allclose(r["avg"], want_avg, atol=1e-6)
    assert torch.allclose(r["mul"], g @ g[5], rtol=1e-4, atol=1e-5)
    gn = g / g.norm(dim=1, keepdim=True).clamp(min=1e-30)
    cos = gn[100:116] @ gn.t()
    want_sim, want_idx = torch.topk(cos, 7, dim=1)
    assert torch.allclose(r["sim16"], want_sim, atol=2e-5)
    assert (r["idx16"] == want_idx).float().mean() > 0.98          # ties / last-place swaps only
    assert r["idx16"][:, 0].tolist() == list(range(100, 116))


@pytest.mark.parametrize("world,d,n", [(2, 300, 10), (8, 300, 10), (2, 64, 16)])
def test_fused_multi_many_negatives(world, d, n, tmp_path):
    """More than 7 negatives per pair (BASELINE.json config 4: 80 M x 300, neg 10): split descriptors through the
    production column-shard kernel (csrc/pairgen.cu, csrc/sgns_pairs.cu) against the dense oracle."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), d, str(tmp_path), n), nprocs=world, join=True)
    r = torch.load(os.path.join(tmp_path, "result.pt"))
    assert [int(x) for x in r["stats"][:, 0]] == r["pairs"]
    upd_ref = r["ref0"] - r["start0"]
    upd_got = r["got0"] - r["start0"]
    assert (upd_got - upd_ref).norm() / upd_ref.norm() < 3e-2


@pytest.mark.parametrize("world,d,mode", [(2, 128, "safe"), (2, 128, "auto"), (8, 512, "safe")])
def test_pair_exchange_formats(world, d, mode, tmp_path, monkeypatch):
    """GW2V_XCHG=safe: 64-bit (value, tag) words instead of 16-byte {f, f, f, tag} chunks (csrc/sgns_pairs.cu);
    auto: the start-up self-test streams 16-byte chunks between all peers and must not see a torn read on this box."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    monkeypatch.setenv("GW2V_XCHG", mode)
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), d, str(tmp_path)), nprocs=world, join=True)
    r = torch.load(os.path.join(tmp_path, "result.pt"))
    assert [int(x) for x in r["stats"][:, 0]] == r["pairs"]
    upd_ref = r["ref0"] - r["start0"]
    upd_got = r["got0"] - r["start0"]
    assert (upd_got - upd_ref).norm() / upd_ref.norm() < 3e-2
    if mode == "safe":
        assert r["debug"] & 32 and r["xchg"] is None
    else:
        assert r["xchg"]["torn"] == 0 and r["xchg"]["observed"] > 0 and not (r["debug"] & 32)


def _worker_tile(rank, world, port, d, nn, out_dir):
    import torch.distributed as dist
    from glint_word2vec_b200.data.sampler import zipf_counts
    from glint_word2vec_b200.models import sgns
    from glint_word2vec_b200.models.engine import EngineOptions, ShardEngine
    from glint_word2vec_b200.models.sgns import SGNSConfig
    from glint_word2vec_b200.parallel.comm import TorchDistComm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        v = 200000
        cfg = SGNSConfig(v, d, 5, 5, seed=11, neg_sharing="tile", tile_negatives=nn)
        eng = ShardEngine(cfg, comm=TorchDistComm(), device=dev, options=EngineOptions(subsample_mode="reference", hot_row_cap=0))
        eng.init_weights()
        eng.set_noise(zipf_counts(v, 10 ** 7, 0.6))
        full1 = (torch.rand(v, eng.shard.padded_vector_size, generator=torch.Generator().manual_seed(5)) - 0.5) * 0.5
        full1[:, d:] = 0
        eng.syn1 = full1[:, rank * eng.shard.cols:(rank + 1) * eng.shard.cols].contiguous().to(dev)
        eng.syn0 = (eng.syn0 * 20.0).contiguous()
        rng = np.random.default_rng(3)
        steps = []
        for t in (5000, 128 * 148 + 77, 300):
            tokens = rng.choice(v, size=t, replace=False).astype(np.int32)
            steps.append((tokens, (np.arange(t) // 41).astype(np.int32)))
        start0 = eng.pull(torch.arange(v)).cpu()
        stats, pos = [], 0
        for tokens, sid in steps:
            stats.append(eng.train_step(tokens, sid, pos, 0, 0.02).cpu())
            pos += len(tokens)
        torch.cuda.synchronize(dev)
        got0 = eng.pull(torch.arange(v)).cpu()
        err = int(eng._cuda._xchg["err"].item())
        if rank == 0:
            ref0, _ = sgns.init_embeddings(v, d, 11)
            ref0 = ref0 * 20.0
            ref1 = full1[:, :d].clone()
            pairs, losses, pos = [], [], 0
            for tokens, sid in steps:
                st = sgns.sgns_minibatch_reference(ref0, ref1, cfg, eng.alias, tokens, sid, pos, 0, 0.02)
                pairs.append(st.pairs)
                losses.append(st.loss)
                pos += len(tokens)
            torch.save({"got0": got0, "ref0": ref0, "start0": start0, "pairs": pairs, "losses": losses,
                        "stats": torch.stack(stats), "err": err}, os.path.join(out_dir, "result_tile.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,d,nn", [(2, 128, 32), (2, 100, 64), (4, 256, 32), (8, 512, 32), (8, 512, 64)])
def test_tile_kernel_column_shards_match_oracle(world, d, nn, tmp_path):
    """neg_sharing="tile" over column shards: the tcgen05 kernel with the in-kernel NVLink exchange of the partial
    S entries (csrc/sgns_tile.cu) against the dense single-process oracle."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker_tile, args=(world, _free_port(), d, nn, str(tmp_path)), nprocs=world, join=True)
    r = torch.load(os.path.join(tmp_path, "result_tile.pt"))
    assert r["err"] == 0
    assert [int(x) for x in r["stats"][:, 0]] == r["pairs"]
    for got, want in zip(r["stats"][:, 1].tolist(), r["losses"]):
        assert abs(got - want) / want < 5e-3
    upd_ref = r["ref0"] - r["start0"]
    upd_got = r["got0"] - r["start0"]
    assert upd_ref.norm() > 0
    assert (upd_got - upd_ref).norm() / upd_ref.norm() < 3e-2


def _worker_nn(rank, world, port, d, out_dir):
    import torch.distributed as dist
    from glint_word2vec_b200.models.engine import EngineOptions, ShardEngine
    from glint_word2vec_b200.models.sgns import SGNSConfig
    from glint_word2vec_b200.parallel.comm import TorchDistComm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        v = 140000 * world + 77
        eng = ShardEngine(SGNSConfig(v, d, 5, 5, seed=11), comm=TorchDistComm(), device=dev,
                          options=EngineOptions(hot_row_cap=0))
        eng.init_weights()
        eng.syn0 = (eng.syn0 * d).contiguous()
        eng._touch()
        full = eng.pull(torch.arange(0, v, 997)).cpu()                 # sanity rows through the column path
        g = torch.Generator().manual_seed(4)
        qs = torch.randn(40, d, generator=g)
        qs[:8] = full[:8] + 0.1 * torch.randn(8, d, generator=g)
        idx, sim = eng.top_k(qs, 10)
        nn = eng._cuda.nn_index()
        used = nn.version == eng._version and nn.overflows == 0
        # dense path as the oracle (GW2V_NN_SELECT=0 switches the select path off)
        os.environ["GW2V_NN_SELECT"] = "0"
        idx0, sim0 = eng.top_k(qs, 10)
        if rank == 0:
            torch.save({"idx": idx, "sim": sim, "idx0": idx0, "sim0": sim0, "used": used},
                       os.path.join(out_dir, "nn.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,d", [(2, 128), (2, 300), (4, 512), (8, 300)])
def test_nn_row_shard_replica_matches_dense_path(world, d, tmp_path):
    """findSynonyms over column shards: the row-sharded serving replica + fused select kernel (ops/nn.py) returns
    the same neighbours as the dense reduce-scatter path."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker_nn, args=(world, _free_port(), d, str(tmp_path)), nprocs=world, join=True)
    r = torch.load(os.path.join(tmp_path, "nn.pt"))
    assert r["used"]
    assert torch.allclose(r["sim"], r["sim0"], atol=2e-5)
    assert (r["idx"] == r["idx0"]).float().mean() > 0.97
