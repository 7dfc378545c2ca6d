"""Oracle + distributed-CPU tiers: the column-sharded algebra must equal the dense
single-process oracle (BASELINE.json config 1: vocab 10k, dim 64,
neg 5, window 5 on CPU world_size=2 over Gloo)."""
import os
import socket
import time

import numpy as np
import pytest
import torch

from glint_word2vec_b200.data.corpus import EncodedCorpus
from glint_word2vec_b200.data.sampler import build_alias, zipf_counts, zipf_tokens
from glint_word2vec_b200.models import matrix_io, sgns, trainer
from glint_word2vec_b200.models.engine import EngineOptions, ShardEngine
from glint_word2vec_b200.models.sgns import SGNSConfig
from glint_word2vec_b200.parallel.comm import Comm
from glint_word2vec_b200.parallel.sharding import make_shard, shard_cols


class FakeComm(Comm):
    """A shard of a world whose collectives are driven by the test."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world


def test_shard_layout():
    assert shard_cols(512, 8) == 64 and shard_cols(300, 8) == 40 and shard_cols(100, 2) == 52
    sh = [make_shard(300, 8, r) for r in range(8)]
    assert sum(s.real_cols for s in sh) == 300 and sh[7].real_cols == 20 and sh[7].col_start == 280
    assert all(s.cols % 4 == 0 for s in sh)
    s5 = [make_shard(100, 5, r) for r in range(5)]                 # the reference default of 5 servers
    assert [s.real_cols for s in s5] == [20] * 5


def test_init_is_shard_invariant():
    cfg = SGNSConfig(300, 100, seed=3)
    full = ShardEngine(cfg, device=torch.device("cpu"))
    full.init_weights()
    parts = []
    for r in range(3):
        e = ShardEngine(cfg, comm=FakeComm(r, 3), device=torch.device("cpu"))
        e.init_weights()
        parts.append(e.syn0[:, :e.shard.real_cols])
    assert torch.equal(torch.cat(parts, 1), full.syn0[:, :100])
    assert float(full.syn0.abs().max()) <= 0.5 / 100 and float(full.syn1.abs().sum()) == 0.0


def test_manual_two_shard_minibatch_equals_oracle():
    """dotprod partials summed across shards + adjust per shard == dense oracle."""
    v, d = 500, 64
    cfg = SGNSConfig(v, d, 5, 5, seed=5)
    counts = zipf_counts(v, 10 ** 5)
    engines = [ShardEngine(cfg, comm=FakeComm(r, 2), device=torch.device("cpu")) for r in range(2)]
    for e in engines:
        e.init_weights()
        e.set_noise(counts)
        e.syn1 = torch.randn(v, e.shard.cols) * 0.1
        e.syn0 = e.syn0 * 30
    ref0 = torch.cat([e.syn0 for e in engines], 1).clone()
    ref1 = torch.cat([e.syn1 for e in engines], 1).clone()
    rng = np.random.default_rng(0)
    tokens = rng.integers(0, v, 400).astype(np.int32)
    sid = (np.arange(400) // 25).astype(np.int32)
    st = sgns.sgns_minibatch_reference(ref0, ref1, cfg, engines[0].alias, tokens, sid, 77, 0, 0.05)
    ci, cj, slot = sgns.enumerate_pairs(cfg, tokens, sid, 77, 0)
    neg = sgns.draw_negatives(cfg, engines[0].alias, np.uint64(77) + ci.astype(np.uint64), slot, 0)
    w = torch.from_numpy(tokens.astype(np.int64)[ci])
    c = torch.from_numpy(tokens.astype(np.int64)[cj])
    ng = torch.from_numpy(neg.astype(np.int64))
    f = sum(e.partial_dots(w, c, ng) for e in engines)             # the all-reduce
    mask = (ng != c[:, None]).float()
    gp = sgns.sigmoid_coeff(f[:, 0], 1.0, 0.05)
    gm = sgns.sigmoid_coeff(f[:, 1:], 0.0, 0.05) * mask
    for e in engines:
        e.adjust(w, c, ng, gp, gm)
    got0 = torch.cat([e.syn0 for e in engines], 1)
    got1 = torch.cat([e.syn1 for e in engines], 1)
    assert st.pairs == w.shape[0] > 0
    assert torch.allclose(got0, ref0, atol=1e-5) and torch.allclose(got1, ref1, atol=1e-5)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from glint_word2vec_b200.parallel.comm import TorchDistComm, init_process_group
    torch.set_num_threads(2)
    init_process_group("gloo", rank=rank, world=world, master_port=port)
    try:
        v, d = 10000, 64                                            # BASELINE.json config 1
        cfg = SGNSConfig(v, d, 5, 5, seed=9)
        counts = zipf_counts(v, 10 ** 6)
        eng = ShardEngine(cfg, comm=TorchDistComm(), device=torch.device("cpu"),
                          options=EngineOptions(batch_size=200, subsample_ratio=1e-3))
        eng.init_weights()
        eng.set_noise(counts)
        alias = build_alias(counts.astype(np.float64))
        toks = zipf_tokens(alias, 6000, seed=4)
        corpus = EncodedCorpus(toks, np.arange(0, 6001, 40, dtype=np.int64))
        rep = trainer.train(eng, corpus, 0.025, 2, train_words=6000)
        vecs = eng.pull(torch.arange(v))
        nrm = eng.norms()
        idx, sim = eng.top_k(vecs[:3], 4)
        avg = eng.pull_average(torch.tensor([1, 2, 3, 9]), torch.tensor([0, 3, 3, 4]))
        matrix_io.save_matrix(eng, os.path.join(out_dir, "m2"))
        if rank == 0:
            torch.save({"vecs": vecs, "pairs": rep.pairs, "loss": rep.loss_per_pair, "nrm": nrm, "idx": idx,
                        "sim": sim, "avg": avg}, os.path.join(out_dir, "r.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gloo_world2_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_gloo_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r = torch.load(os.path.join(tmp_path, "r.pt"))
    v, d = 10000, 64
    cfg = SGNSConfig(v, d, 5, 5, seed=9)
    counts = zipf_counts(v, 10 ** 6)
    eng = ShardEngine(cfg, device=torch.device("cpu"), options=EngineOptions(batch_size=200, subsample_ratio=1e-3))
    eng.init_weights()
    eng.set_noise(counts)
    toks = zipf_tokens(build_alias(counts.astype(np.float64)), 6000, seed=4)
    rep = trainer.train(eng, EncodedCorpus(toks, np.arange(0, 6001, 40, dtype=np.int64)), 0.025, 2, train_words=6000)
    vecs = eng.pull(torch.arange(v))
    assert rep.pairs == r["pairs"] > 0
    assert abs(rep.loss_per_pair - r["loss"]) < 1e-4
    assert torch.allclose(vecs, r["vecs"], atol=1e-5)
    assert torch.allclose(eng.norms(), r["nrm"], atol=1e-5)
    assert r["idx"][:, 0].tolist() == [0, 1, 2]
    ref_avg = torch.stack([vecs[[1, 2, 3]].mean(0), torch.zeros(d), vecs[9]])
    assert torch.allclose(r["avg"], ref_avg, atol=1e-5)
    # re-sharding on load (Q12): saved by 2 shards, loaded by 1 and by 3
    one = matrix_io.load_matrix(os.path.join(tmp_path, "m2"), Comm(), torch.device("cpu"))
    assert torch.allclose(one.syn0[:, :d], vecs, atol=1e-5)
    parts = [matrix_io.load_matrix(os.path.join(tmp_path, "m2"), FakeComm(k, 3), torch.device("cpu")) for k in range(3)]
    cat = torch.cat([p.syn0[:, :p.shard.real_cols] for p in parts], 1)
    assert torch.allclose(cat, vecs, atol=1e-5)
    meta = matrix_io.read_meta(os.path.join(tmp_path, "m2"))
    assert meta["num_shards"] == 2 and [s["cols"] for s in meta["shards"]] == [32, 32]


def test_training_reduces_loss_and_is_deterministic():
    v, d = 2000, 32
    counts = zipf_counts(v, 10 ** 5)
    toks = zipf_tokens(build_alias(counts.astype(np.float64)), 20000, seed=1)
    corpus = EncodedCorpus(toks, np.arange(0, 20001, 50, dtype=np.int64))

    def run():
        eng = ShardEngine(SGNSConfig(v, d, seed=2), device=torch.device("cpu"),
                          options=EngineOptions(batch_size=100, subsample_ratio=1e-2))
        eng.init_weights()
        eng.set_noise(counts)
        rep = trainer.train(eng, corpus, 0.05, 3, train_words=20000)
        return eng, rep
    e1, r1 = run()
    e2, r2 = run()
    assert torch.equal(e1.syn0, e2.syn0) and r1.pairs == r2.pairs
    first, last = r1.history[0]["loss_per_pair"], r1.history[-1]["loss_per_pair"]
    assert last < first < 6 * np.log(2) + 0.2
    assert r1.final_alpha < 0.05 and r1.words == 3 * 20000


def test_zero_iterations_leaves_random_init():
    eng = ShardEngine(SGNSConfig(100, 16, seed=1), device=torch.device("cpu"))
    eng.init_weights()
    eng.set_noise(np.ones(100, dtype=np.int64))
    before = eng.syn0.clone()
    rep = trainer.train(eng, EncodedCorpus(np.arange(50, dtype=np.int32), np.array([0, 50])), 0.025, 0)   # Q8
    assert rep.pairs == 0 and torch.equal(eng.syn0, before)


def test_checkpoint_resume_reproduces_uninterrupted_run(tmp_path):
    """Kill-and-resume == uninterrupted run (bitwise on CPU), also across a shard-count change."""
    from glint_word2vec_b200.models import checkpoint
    v, d = 1500, 32
    counts = zipf_counts(v, 10 ** 5)
    toks = zipf_tokens(build_alias(counts.astype(np.float64)), 12000, seed=8)
    corpus = EncodedCorpus(toks, np.arange(0, 12001, 40, dtype=np.int64))

    def fresh():
        e = ShardEngine(SGNSConfig(v, d, seed=4), device=torch.device("cpu"),
                        options=EngineOptions(batch_size=100, step_tokens=1000))
        e.init_weights()
        e.set_noise(counts)
        return e
    ref = fresh()
    trainer.train(ref, corpus, 0.05, 2, train_words=12000)
    # interrupted run: checkpoint every 5 steps, "crash" after 17 steps
    run = fresh()
    ck = checkpoint.Checkpointer(run, str(tmp_path), 5, dict(learning_rate=0.05, num_iterations=2,
                                                             train_words=12000, step_tokens=1000))
    class Crash(Exception):
        pass
    n = {"steps": 0}

    def fn(k, s):
        ck(k, s)
        n["steps"] += 1
        if n["steps"] == 17:
            raise Crash()
    with pytest.raises(Crash):
        trainer.train(run, corpus, 0.05, 2, train_words=12000, checkpoint_fn=fn)
    path = checkpoint.latest(str(tmp_path))
    assert path is not None and path.endswith("ckpt-0001-00000003")      # 12 steps/iteration -> step 15 = (1, 3)
    eng, rep = checkpoint.resume(str(tmp_path), corpus, counts, Comm(), torch.device("cpu"),
                                 EngineOptions(batch_size=100))
    assert torch.equal(eng.syn0, ref.syn0) and torch.equal(eng.syn1, ref.syn1)
    assert rep.steps == 24 - 15


def test_killed_shard_aborts_cleanly_and_training_resumes(tmp_path):
    """Fault injection: kill one shard process of an integrated group in the middle of `fit`.
    The client's request must FAIL (not hang), the group must stop, and a fresh group must be able to resume
    from the last checkpoint and finish the run."""
    import threading
    import psutil
    from glint_word2vec_b200.models import checkpoint
    from glint_word2vec_b200.parallel import cluster
    v, d = 1500, 32
    counts = zipf_counts(v, 10 ** 5)
    toks = zipf_tokens(build_alias(counts.astype(np.float64)), 60000, seed=8)
    corpus = EncodedCorpus(toks, np.arange(0, 60001, 40, dtype=np.int64))
    cfg = SGNSConfig(v, d, seed=4)
    opts = {"batch_size": 100, "step_tokens": 500}
    ckdir = str(tmp_path / "ck")
    train_opts = {"checkpoint_dir": ckdir, "checkpoint_every_steps": 5}

    h = cluster.spawn_integrated(2, "cpu", opts)
    launcher = h._procs[0]
    try:
        h.create(cfg, opts, counts)
        err = {}

        def run():
            try:
                h.fit(corpus, 0.05, 3, 60000, train_opts=train_opts)
            except Exception as e:          # ServerError / EOFError / ConnectionError: the request failed
                err["e"] = e
        th = threading.Thread(target=run)
        th.start()
        t0 = time.time()
        while checkpoint.latest(ckdir) is None:                       # training is under way
            assert time.time() - t0 < 120 and th.is_alive()
            time.sleep(0.05)
        ranks = [c for c in psutil.Process(launcher.pid).children() if "--rank" in c.cmdline()]
        victim = [c for c in ranks if c.cmdline()[c.cmdline().index("--rank") + 1] == "1"][0]
        victim.kill()                                                 # exact pid of the shard we started
        th.join(timeout=120)
        assert not th.is_alive(), "fit() hung after a shard died"
        assert "e" in err, "fit() must fail when a shard dies"
        assert launcher.wait(timeout=60) != 0                         # the launcher stopped the whole group
    finally:
        if launcher.poll() is None:
            launcher.kill()
    st_path = checkpoint.latest(ckdir)
    assert st_path is not None
    # recovery: fresh group, resume from the last checkpoint
    h2 = cluster.spawn_integrated(2, "cpu", opts)
    try:
        h2.create(cfg, opts, counts)
        rep = h2.fit(corpus, 0.05, 3, 60000, train_opts=dict(train_opts, resume=True))
        from glint_word2vec_b200.data.corpus import iter_steps
        total_steps = 3 * sum(1 for _ in iter_steps(corpus, 500))
        assert 0 < rep["steps"] < total_steps                          # only the remainder was trained
        vecs = h2.pull(np.arange(10))
        assert np.isfinite(vecs).all()
    finally:
        h2.terminate()


def test_tile_shared_negatives_engine_equals_oracle_and_shards_agree():
    """neg_sharing="tile" (docs/round2_tile_gemm.md): the un-fused GEMM path of the engine == the dense oracle,
    for one shard and for two column shards (partials summed), and the negative weights keep n negatives per
    pair in expectation."""
    v, d = 800, 48
    cfg = SGNSConfig(v, d, 5, 5, seed=5, neg_sharing="tile", tile_centres=32, tile_negatives=16)
    counts = zipf_counts(v, 10 ** 5)
    rng = np.random.default_rng(0)
    tokens = rng.integers(0, v, 300).astype(np.int32)
    sid = (np.arange(300) // 25).astype(np.int32)

    def make(world, rank):
        e = ShardEngine(cfg, comm=FakeComm(rank, world), device=torch.device("cpu"),
                        options=EngineOptions(batch_size=300, hot_row_cap=0))
        e.init_weights()
        e.set_noise(counts)
        g = torch.Generator().manual_seed(1)
        full1 = torch.randn(v, e.shard.padded_vector_size, generator=g) * 0.1
        full1[:, d:] = 0
        e.syn1 = full1[:, rank * e.shard.cols:(rank + 1) * e.shard.cols].contiguous()
        e.syn0 = e.syn0 * 30
        return e
    one = make(1, 0)
    ref0, ref1 = one.syn0[:, :d].clone(), one.syn1[:, :d].clone()
    st = sgns.sgns_minibatch_reference(ref0, ref1, cfg, one.alias, tokens, sid, 77, 0, 0.05)
    got = one.train_step(tokens, sid, 77, 0, 0.05)
    assert int(got[0]) == st.pairs > 0
    assert abs(float(got[1]) - st.loss) / st.loss < 1e-5
    assert torch.allclose(one.syn0[:, :d], ref0, atol=1e-5) and torch.allclose(one.syn1[:, :d], ref1, atol=1e-5)
    # weights: sum over a centre's shared negatives = m_i * n
    ci, _, _ = sgns.enumerate_pairs(cfg, tokens, sid, 77, 0)
    centres, m, tile = sgns.tile_terms(cfg, ci)
    assert (tile == centres // 32).all() and m.sum() == len(ci)
    assert np.allclose(m * cfg.negatives / cfg.tile_negatives * cfg.tile_negatives, m * cfg.negatives)
    # every shard regenerates the same shared negatives (pure function of the stream position)
    tn = sgns.tile_negatives(cfg, one.alias, 77, np.unique(tile), 0)
    assert tn.shape == (len(np.unique(tile)), 16)
    assert np.array_equal(tn, sgns.tile_negatives(cfg, one.alias, 77, np.unique(tile), 0))

    # two column shards through a shared in-memory all-reduce
    class PairComm(FakeComm):
        box = {}

        def all_reduce_sum(self, t):
            PairComm.box.setdefault("parts", []).append(t.clone())
            return t
    shards = [make(2, r) for r in range(2)]
    # pass 1 records each shard's partial dots, pass 2 replays the step with their sum as the all-reduce result
    for e in shards:
        e.comm = PairComm(e.comm.rank, 2)
    snap = [(e.syn0.clone(), e.syn1.clone()) for e in shards]
    PairComm.box.clear()
    for e in shards:
        e.train_step(tokens, sid, 77, 0, 0.05)                     # each returns its own partial as "reduced"
    partials = PairComm.box["parts"]
    total = partials[0] + partials[1]

    class SumComm(FakeComm):
        def all_reduce_sum(self, t):
            return total.clone()
    for e, (s0, s1) in zip(shards, snap):
        e.syn0, e.syn1 = s0, s1
        e.comm = SumComm(e.comm.rank, 2)
        e.train_step(tokens, sid, 77, 0, 0.05)
    cat0 = torch.cat([e.syn0 for e in shards], 1)[:, :d]
    cat1 = torch.cat([e.syn1 for e in shards], 1)[:, :d]
    assert torch.allclose(cat0, ref0, atol=1e-5) and torch.allclose(cat1, ref1, atol=1e-5)


def test_server_errors_are_reported_and_the_group_survives():
    """A failing request (unknown op, unknown matrix, bad arguments) comes back as ServerError with the remote
    traceback; the shard-server group keeps serving (the reference's futures time out or are swallowed, Q4)."""
    from glint_word2vec_b200.parallel import cluster
    h = cluster.spawn_integrated(2, "cpu", {"batch_size": 50})
    try:
        with pytest.raises(cluster.ServerError, match="unknown op"):
            h._call("no_such_op")
        with pytest.raises(cluster.ServerError):
            h._call("pull", "no-such-matrix", np.arange(3))
        v, d = 200, 16
        h.create(SGNSConfig(v, d, seed=1), {"batch_size": 50}, zipf_counts(v, 10 ** 4))
        with pytest.raises(cluster.ServerError):
            h.pull(np.array([v + 5]))                                # row out of range
        vec = h.pull(np.array([0, 1]))                               # still alive on both shards
        assert vec.shape == (2, d) and np.isfinite(vec).all()
        assert h._call("info")["world"] == 2
    finally:
        h.terminate()


def test_index_validation_on_the_host():
    """Bad row / token indices are rejected before they can reach a kernel (which does not bounds-check)."""
    v, d = 100, 8
    e = ShardEngine(SGNSConfig(v, d, seed=1), device=torch.device("cpu"))
    e.init_weights()
    with pytest.raises(IndexError):
        e.pull(torch.tensor([0, v]))
    with pytest.raises(IndexError):
        e.pull_average(torch.tensor([-1]), torch.tensor([0, 1]))
    with pytest.raises(ValueError):
        e.pull_average(torch.tensor([1, 2]), torch.tensor([0, 3]))
    from glint_word2vec_b200 import ServerSideGlintWord2Vec
    est = ServerSideGlintWord2Vec(inputCol="s", outputCol="v", vectorSize=8, numParameterServers=1,
                                  parameterServerConfig={"device": "cpu"})
    with pytest.raises(ValueError):
        est.fitEncoded(np.array([0, 1, 7], dtype=np.int32), np.array([0, 3]), np.array([5, 4, 3]))


def test_streamed_matrix_io_round_trip_in_small_chunks(tmp_path, monkeypatch):
    """Shards are written and read in row chunks (never a full host copy); tiny chunk size forces the loop."""
    monkeypatch.setattr(matrix_io, "CHUNK_BYTES", 4096)
    v, d = 777, 20
    e = ShardEngine(SGNSConfig(v, d, seed=2), device=torch.device("cpu"))
    e.init_weights()
    e.syn1 = torch.randn(v, e.shard.cols)
    e.syn1[:, d:] = 0
    matrix_io.save_matrix(e, str(tmp_path / "m"))
    raw = np.load(str(tmp_path / "m" / "matrix" / "syn0.00of01.npy"))
    assert raw.shape == (v, d) and np.array_equal(raw, e.syn0[:, :d].numpy())
    back = matrix_io.load_matrix(str(tmp_path / "m"), Comm(), torch.device("cpu"))
    assert torch.equal(back.syn0, e.syn0) and torch.equal(back.syn1, e.syn1)


def test_checkpoint_pruning_never_deletes_latest_and_stale_runs_are_refused(tmp_path):
    """pruning sorted every ckpt-* by name, so in a directory holding an older run's ckpt-0002-* the
    checkpoint just written was deleted right after LATEST was pointed at it; and resume=true picked up a finished
    run's state.  Now: only own checkpoints are pruned, never LATEST; a fresh run refuses a used directory; resume
    checks a corpus / configuration fingerprint."""
    from glint_word2vec_b200.data.corpus import EncodedCorpus
    from glint_word2vec_b200.models import checkpoint
    from glint_word2vec_b200.models.engine import EngineOptions, ShardEngine
    from glint_word2vec_b200.models.sgns import SGNSConfig
    rng = np.random.default_rng(0)
    v = 200

    def make(seed=3):
        eng = ShardEngine(SGNSConfig(v, 16, seed=seed), device=torch.device("cpu"),
                          options=EngineOptions(subsample_mode="reference", step_tokens=500))
        eng.init_weights()
        eng.set_noise(np.arange(v, 0, -1))
        return eng
    toks = rng.integers(0, v, size=4000).astype(np.int32)
    corpus = EncodedCorpus(toks, np.arange(0, 4001, 40, dtype=np.int64))
    d = str(tmp_path / "ck")
    os.makedirs(os.path.join(d, "ckpt-0002-00000009", "matrix"))                  # leftovers of an earlier run
    eng = make()
    hyper = dict(learning_rate=0.05, num_iterations=1, train_words=4000, step_tokens=500)
    ck = checkpoint.Checkpointer(eng, d, 1, hyper, keep=2, fp=checkpoint.fingerprint(eng, corpus))
    for step in (1, 2, 3, 4):
        ck.save(0, step)
        assert checkpoint.latest(d) == os.path.join(d, f"ckpt-0000-{step:08d}")   # the newest one always survives
    left = sorted(x for x in os.listdir(d) if x.startswith("ckpt-"))
    assert left == ["ckpt-0000-00000003", "ckpt-0000-00000004", "ckpt-0002-00000009"]      # foreign directory untouched
    # a fresh run refuses the used directory ...
    with pytest.raises(FileExistsError):
        checkpoint.train_with_checkpoints(make(), corpus, 0.05, 1, 4000, d, 2)
    # ... unless told to overwrite it
    checkpoint.train_with_checkpoints(make(), corpus, 0.05, 1, 4000, d, 2, overwrite=True)
    assert "ckpt-0002-00000009" not in os.listdir(d) and checkpoint.latest(d) is not None
    # resume with another corpus (or seed) is an error, not a silent continuation of somebody else's run
    other = EncodedCorpus(rng.integers(0, v, size=4000).astype(np.int32), corpus.offsets)
    from glint_word2vec_b200.parallel.comm import Comm
    with pytest.raises(ValueError, match="does not belong to this run"):
        checkpoint.resume(d, other, np.arange(v, 0, -1), Comm(), torch.device("cpu"),
                          EngineOptions(subsample_mode="reference", hot_row_cap=0))


def test_step_prefetch_keeps_order_and_propagates_errors():
    """models/trainer.py::_prefetch: the background producer hands items over in order, re-raises its exception in
    the consumer, and stops when the consumer walks away."""
    from glint_word2vec_b200.models.trainer import _prefetch
    assert list(_prefetch(iter(range(100)), 4)) == list(range(100))

    def boom():
        yield 1
        yield 2
        raise RuntimeError("producer failed")
    got = []
    with pytest.raises(RuntimeError, match="producer failed"):
        for x in _prefetch(boom(), 2):
            got.append(x)
    assert got == [1, 2]
    produced = []

    def slow():
        for i in range(1000):
            produced.append(i)
            yield i
    g = _prefetch(slow(), 2)
    assert next(g) == 0
    g.close()                                  # abandoned consumer: the producer must stop within its hand-over timeout
    import time
    time.sleep(0.6)
    n = len(produced)
    time.sleep(0.4)
    assert len(produced) == n and n < 20


def test_training_report_carries_phase_times(tmp_path):
    """StepTimer is wired into trainer.train: cumulative per-phase times in every metrics record and in the report."""
    import json
    from glint_word2vec_b200.data.synthetic import synthetic_capitals_corpus
    from glint_word2vec_b200 import ServerSideGlintWord2Vec
    mp = tmp_path / "metrics.jsonl"
    m = ServerSideGlintWord2Vec(inputCol="s", outputCol="v", vectorSize=8, minCount=5, seed=1, numParameterServers=1,
                                parameterServerConfig={"device": "cpu", "metrics_path": str(mp)}).fit(synthetic_capitals_corpus()[:400])
    try:
        rep = m.trainingReport
        assert rep["device_ms"]["sgns_step"]["n"] == rep["steps"] and rep["device_ms"]["sgns_step"]["ms"] > 0
        recs = [json.loads(l) for l in mp.read_text().splitlines()]
        assert recs and "device_ms" in recs[-1] and "pairs_per_sec" in recs[-1]
    finally:
        m.stop()
