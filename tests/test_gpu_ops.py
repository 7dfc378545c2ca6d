"""GPU numerics: every sm_100a kernel against a plain PyTorch / numpy fp32
reference of the same op (SURVEY.md 4.3 "GPU kernels" tier)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from glint_word2vec_b200.data.sampler import (build_alias, keep_thresholds, zipf_counts,
                                              zipf_tokens)
from glint_word2vec_b200.models import sgns
from glint_word2vec_b200.models.engine import EngineOptions, ShardEngine
from glint_word2vec_b200.models.sgns import SGNSConfig


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def _C():
    from glint_word2vec_b200.ops.cuda import extension
    return extension()


def test_extension_is_native():
    _dev()
    C = _C()
    assert C.__file__.endswith(".so") and "glint_word2vec_b200" in C.__file__


def test_zipf_stream_matches_numpy_philox():
    dev = _dev()
    counts = zipf_counts(5000, 10 ** 6)
    alias = build_alias(counts.astype(np.float64))
    ref = zipf_tokens(alias, 4096, seed=1234567890123, pos0=(1 << 33) + 5)
    out = torch.empty(4096, dtype=torch.int32, device=dev)
    _C().zipf_stream(torch.from_numpy(alias.packed()).to(dev), 1234567890123, (1 << 33) + 5, out)
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("world,rank,d", [(1, 0, 100), (2, 1, 100), (8, 7, 300)])
def test_init_syn0_matches_oracle(world, rank, d):
    dev = _dev()
    from glint_word2vec_b200.parallel.sharding import make_shard
    sh = make_shard(d, world, rank)
    v = 777
    syn0 = torch.empty(v, sh.cols, device=dev)
    _C().init_syn0(syn0, rank * sh.cols, d, 42)
    full, _ = sgns.init_embeddings(v, sh.padded_vector_size, 42, scale_dim=d)
    full[:, d:] = 0
    ref = full[:, rank * sh.cols:(rank + 1) * sh.cols]
    assert torch.equal(syn0.cpu(), ref)


def test_subsample_compact_matches_numpy():
    dev = _dev()
    C = _C()
    v = 2000
    counts = zipf_counts(v, 10 ** 6)
    keep = keep_thresholds(counts, 1e-3, "word2vec")
    alias = build_alias(counts.astype(np.float64))
    for t in (1, 777, 2048, 50000):
        tokens = zipf_tokens(alias, t, seed=3)
        sid = (np.arange(t) // 17).astype(np.int32)
        mask = sgns.subsample_mask(tokens, keep, seed=99, iteration=2, raw_pos0=1000)
        tok_out = torch.zeros(t, dtype=torch.int32, device=dev)
        sid_out = torch.zeros(t, dtype=torch.int32, device=dev)
        count = torch.zeros(1, dtype=torch.int32, device=dev)
        tiles = torch.zeros(int(C.subsample_max_blocks(t)) + 1, dtype=torch.int32, device=dev)
        for _ in (1, 2):          # second launch re-uses the tile workspace
            C.subsample_compact(torch.from_numpy(tokens).to(dev), torch.from_numpy(sid).to(dev), t,
                                torch.from_numpy(keep.view(np.int32).copy()).to(dev), 99, 2, 1000,
                                tok_out, sid_out, count, tiles)
            n = int(count.item())
            assert n == int(mask.sum())
            assert np.array_equal(tok_out[:n].cpu().numpy(), tokens[mask])
            assert np.array_equal(sid_out[:n].cpu().numpy(), sid[mask])


def _make_engine(dev, v, d, window=5, n=5, window_mode="reference", seed=7):
    cfg = SGNSConfig(v, d, window, n, seed=seed, window_mode=window_mode)
    eng = ShardEngine(cfg, device=dev, options=EngineOptions(subsample_mode="reference", hot_row_cap=0))
    eng.init_weights()
    counts = zipf_counts(v, 10 ** 7, 0.6)
    eng.set_noise(counts)
    return eng, counts


@pytest.mark.parametrize("d,window,n,wmode", [(64, 5, 5, "reference"), (100, 5, 5, "reference"),
                                              (128, 3, 7, "word2vec_c"), (512, 5, 5, "reference"),
                                              (300, 5, 10, "word2vec_c"), (40, 5, 5, "reference"),
                                              (64, 4, 16, "reference"), (128, 5, 21, "reference")])
def test_sgns_step_single_matches_oracle(d, window, n, wmode):
    """Distinct centre/context tokens and V >> negatives: concurrent warps almost never
    re-read a row another warp has just updated, so the Hogwild kernel must match the
    summed mini-batch oracle closely (sequential-vs-batch oracle runs differ by < 1e-2 here)."""
    dev = _dev()
    v = 200000
    eng, counts = _make_engine(dev, v, d, window, n, wmode)
    g = torch.Generator().manual_seed(0)
    # |u| = |v| = 0.5 for every d, so one pair's update is ~0.25 % of a row at alpha = 0.002
    syn1 = torch.randn(v, eng.shard.cols, generator=g) * (0.5 / d ** 0.5)
    syn1[:, d:] = 0
    eng.syn1 = syn1.to(dev)
    syn0 = torch.randn(v, eng.shard.cols, generator=g) * (0.5 / d ** 0.5)
    syn0[:, d:] = 0
    eng.syn0 = syn0.to(dev)
    t = 3000
    rng = np.random.default_rng(1)
    tokens = rng.choice(v, size=t, replace=False).astype(np.int32)
    sid = (np.arange(t) // 37).astype(np.int32)
    ref0, ref1 = syn0[:, :d].clone(), syn1[:, :d].clone()
    cfg = eng.cfg
    # small alpha: the kernels apply du per pair (word2vec.c order) while the oracle sums the whole
    # mini-batch from pre-update rows; the difference is second order in alpha
    st = sgns.sgns_minibatch_reference(ref0, ref1, cfg, eng.alias, tokens, sid, 12345, 1, 0.002)
    stats = eng.train_step(tokens, sid, 12345, 1, 0.002).cpu()
    assert int(stats[0]) == st.pairs
    assert int(stats[3]) == t
    assert abs(float(stats[1]) - st.loss) / st.loss < 2e-3
    got0, got1 = eng.syn0.cpu()[:, :d], eng.syn1.cpu()[:, :d]
    d0, d1 = got0 - syn0[:, :d], got1 - syn1[:, :d]
    r0, r1 = ref0 - syn0[:, :d], ref1 - syn1[:, :d]
    assert (r0.abs().sum() > 0) and (r1.abs().sum() > 0)
    assert (d0 - r0).norm() / r0.norm() < 2e-2
    assert (d1 - r1).norm() / r1.norm() < 2e-2
    # padding columns never move
    assert float(eng.syn0[:, d:].abs().sum()) == 0.0 or eng.shard.cols == d


def test_sgns_step_sigmoid_table_mode():
    """sigmoid_mode="table": the kernels look sigma up in the reference's 1000-entry table (MLLIB:281-302,
    index scale 83.0) - diffed against the oracle in the same mode, with dots spread over [-6, 6] and beyond."""
    dev = _dev()
    v, d = 50000, 64
    cfg = SGNSConfig(v, d, 5, 5, seed=7, sigmoid_mode="table")
    eng = ShardEngine(cfg, device=dev, options=EngineOptions(subsample_mode="reference", hot_row_cap=0))
    eng.init_weights()
    eng.set_noise(zipf_counts(v, 10 ** 7, 0.6))
    g = torch.Generator().manual_seed(0)
    syn1 = torch.randn(v, d, generator=g) * (2.0 / d ** 0.5)
    syn0 = torch.randn(v, d, generator=g) * (2.0 / d ** 0.5)
    eng.syn0, eng.syn1 = syn0.to(dev), syn1.to(dev)
    t = 2000
    tokens = np.random.default_rng(1).choice(v, size=t, replace=False).astype(np.int32)
    sid = (np.arange(t) // 37).astype(np.int32)
    ref0, ref1 = syn0.clone(), syn1.clone()
    st = sgns.sgns_minibatch_reference(ref0, ref1, cfg, eng.alias, tokens, sid, 0, 0, 0.0005)
    ex0, ex1 = syn0.clone(), syn1.clone()
    import dataclasses
    sgns.sgns_minibatch_reference(ex0, ex1, dataclasses.replace(cfg, sigmoid_mode="exact"), eng.alias, tokens, sid,
                                  0, 0, 0.0005)
    stats = eng.train_step(tokens, sid, 0, 0, 0.0005).cpu()
    assert int(stats[0]) == st.pairs
    d0, r0, e0 = eng.syn0.cpu() - syn0, ref0 - syn0, ex0 - syn0
    err_table = float((d0 - r0).norm() / r0.norm())
    assert err_table < 2e-2
    # the two modes are distinguishable on this input
    assert float((e0 - r0).norm() / r0.norm()) > 1e-4


def test_sgns_step_neg_sharing_centre_matches_oracle():
    """neg_sharing="centre" (negatives drawn once per centre, shared by its pairs) against the oracle in the
    same mode; the pair count and the negatives are bit-identical, the updates agree to Hogwild tolerance."""
    dev = _dev()
    v, d = 200000, 64
    cfg = SGNSConfig(v, d, 5, 5, seed=7, neg_sharing="centre")
    eng = ShardEngine(cfg, device=dev, options=EngineOptions(subsample_mode="reference", hot_row_cap=0))
    eng.init_weights()
    eng.set_noise(zipf_counts(v, 10 ** 7, 0.6))
    g = torch.Generator().manual_seed(0)
    syn1 = torch.randn(v, d, generator=g) * (0.5 / d ** 0.5)
    syn0 = torch.randn(v, d, generator=g) * (0.5 / d ** 0.5)
    eng.syn0, eng.syn1 = syn0.to(dev), syn1.to(dev)
    t = 3000
    tokens = np.random.default_rng(1).choice(v, size=t, replace=False).astype(np.int32)
    sid = (np.arange(t) // 37).astype(np.int32)
    ref0, ref1 = syn0.clone(), syn1.clone()
    st = sgns.sgns_minibatch_reference(ref0, ref1, cfg, eng.alias, tokens, sid, 4242, 0, 0.002)
    stats = eng.train_step(tokens, sid, 4242, 0, 0.002).cpu()
    assert int(stats[0]) == st.pairs
    assert abs(float(stats[1]) - st.loss) / st.loss < 2e-3
    d0, r0 = eng.syn0.cpu() - syn0, ref0 - syn0
    d1, r1 = eng.syn1.cpu() - syn1, ref1 - syn1
    assert (d0 - r0).norm() / r0.norm() < 2e-2
    assert (d1 - r1).norm() / r1.norm() < 3e-2       # shared negatives: the same row takes several updates per centre


def test_unfused_transport_on_gpu_has_minibatch_semantics():
    """transport="nccl": dotprod -> (library) all-reduce -> adjust with torch ops on the device - the reference's
    exact mini-batch semantics (all dots of a mini-batch from pre-update rows), so it matches the oracle to fp32
    rounding, unlike the Hogwild kernels."""
    dev = _dev()
    v, d = 20000, 48
    cfg = SGNSConfig(v, d, 5, 5, seed=7)
    eng = ShardEngine(cfg, device=dev, options=EngineOptions(subsample_mode="reference", transport="nccl",
                                                             batch_size=4000))
    assert eng.unfused
    eng.init_weights()
    eng.set_noise(zipf_counts(v, 10 ** 6, 0.8))
    g = torch.Generator().manual_seed(0)
    syn1 = torch.randn(v, d, generator=g) * 0.1
    syn0 = torch.randn(v, d, generator=g) * 0.1
    eng.syn0, eng.syn1 = syn0.to(dev), syn1.to(dev)
    t = 4000
    tokens = np.random.default_rng(1).integers(0, v, size=t).astype(np.int32)
    sid = (np.arange(t) // 37).astype(np.int32)
    ref0, ref1 = syn0.clone(), syn1.clone()
    st = sgns.sgns_minibatch_reference(ref0, ref1, cfg, eng.alias, tokens, sid, 5, 0, 0.025)
    stats = torch.as_tensor(eng.train_step(tokens, sid, 5, 0, 0.025)).cpu()
    assert int(stats[0]) == st.pairs
    assert torch.allclose(eng.syn0.cpu(), ref0, atol=2e-6)
    assert torch.allclose(eng.syn1.cpu(), ref1, atol=2e-6)
    with pytest.raises(ValueError):
        EngineOptions(transport="carrier-pigeon")


def test_sgns_step_hot_rows_hogwild_close():
    """Dense collisions (tiny vocabulary): updates race by design; the result
    must still be close to the summed mini-batch oracle and finite."""
    dev = _dev()
    v, d = 50, 64
    eng, _ = _make_engine(dev, v, d)
    t = 4000
    rng = np.random.default_rng(2)
    tokens = rng.integers(0, v, size=t).astype(np.int32)
    sid = np.zeros(t, dtype=np.int32)
    ref0, ref1 = eng.syn0.cpu().clone(), eng.syn1.cpu().clone()
    st = sgns.sgns_minibatch_reference(ref0, ref1, eng.cfg, eng.alias, tokens, sid, 0, 0, 0.01)
    stats = eng.train_step(tokens, sid, 0, 0, 0.01).cpu()
    assert int(stats[0]) == st.pairs
    assert torch.isfinite(eng.syn0).all() and torch.isfinite(eng.syn1).all()
    # racing updates: same direction as the summed mini-batch update, comparable magnitude
    upd, ref_upd = eng.syn1.cpu().flatten(), ref1.flatten()          # syn1 starts at zero
    cos = float(torch.dot(upd, ref_upd) / (upd.norm() * ref_upd.norm()))
    assert cos > 0.5 and 0.1 < float(upd.norm() / ref_upd.norm()) < 10.0


def test_async_steps_stage_inputs_safely():
    """The host queues many steps ahead of the device (trainer.train does): every step must train on ITS OWN
    tokens - pinned staging buffers are a ring guarded by events, never overwritten while an H2D is pending.
    Pair counts are a pure function of (positions, sentence ids), so a clobbered buffer shows up exactly."""
    dev = _dev()
    v, d = 50000, 64
    eng, _ = _make_engine(dev, v, d)
    rng = np.random.default_rng(5)
    steps = []
    for s in range(40):
        t = int(rng.integers(2000, 30000))
        tokens = rng.integers(0, v, size=t).astype(np.int32)
        sid = (np.arange(t) // int(rng.integers(3, 60))).astype(np.int32)
        steps.append((tokens, sid, 1000 * s))
    handles = [eng.train_step_async(tok, sid, pos0, 0, 0.001) for tok, sid, pos0 in steps]     # no sync in between
    for (tok, sid, pos0), h in zip(steps, handles):
        ci, _, _ = sgns.enumerate_pairs(eng.cfg, tok, sid, pos0, 0)
        st = h.result()
        assert int(st[0]) == len(ci)
        assert int(st[3]) == len(tok)


def test_unsupported_shape_is_rejected_loudly():
    """No silent fall-back kernels: shapes outside the fused step's limits raise with the limit named."""
    dev = _dev()
    with pytest.raises(ValueError, match="negatives <= 21"):
        ShardEngine(SGNSConfig(1000, 64, 5, 30), device=dev)
    with pytest.raises(ValueError, match="window <= 11"):
        ShardEngine(SGNSConfig(1000, 64, 15, 5), device=dev)


def test_zero_pair_step_is_noop():
    dev = _dev()
    eng, _ = _make_engine(dev, 1000, 64)
    before0, before1 = eng.syn0.clone(), eng.syn1.clone()
    tokens = np.arange(10, dtype=np.int32)
    sid = np.arange(10, dtype=np.int32)          # every token its own sentence -> no pairs
    stats = eng.train_step(tokens, sid, 0, 0, 0.025).cpu()
    assert int(stats[0]) == 0
    assert torch.equal(eng.syn0, before0) and torch.equal(eng.syn1, before1)


def test_inference_kernels_match_torch():
    dev = _dev()
    v, d = 30011, 100
    eng, _ = _make_engine(dev, v, d)
    eng.syn0 = (torch.randn(v, eng.shard.cols) * (torch.arange(eng.shard.cols) < d)).to(dev)
    eng.syn0[5] = 0                                        # zero-norm row scores 0 (MLLIB:602-607)
    full = eng.syn0.cpu()[:, :d]
    rows = torch.tensor([0, 5, 17, v - 1, 17])
    assert torch.equal(eng.pull(rows).cpu(), full[rows])
    flat = torch.tensor([1, 2, 3, 10, 11, 7])
    offs = torch.tensor([0, 3, 3, 5, 6])
    avg = eng.pull_average(flat, offs).cpu()
    ref = torch.stack([full[[1, 2, 3]].mean(0), torch.zeros(d), full[[10, 11]].mean(0), full[7]])
    assert torch.allclose(avg, ref, atol=1e-6)
    assert torch.allclose(eng.norms().cpu(), full.norm(dim=1), rtol=1e-5, atol=1e-6)
    q = torch.randn(3, d)
    sc = eng._scores(q).cpu()
    assert torch.allclose(sc, q @ full.t(), rtol=1e-4, atol=1e-4)
    idx, sim = eng.top_k(q, 12)
    qn = q / q.norm(dim=1, keepdim=True)
    nr = full.norm(dim=1)
    cos = (qn @ full.t()) / torch.where(nr > 0, nr, torch.ones_like(nr))
    cos[:, nr == 0] = 0
    rs, ri = torch.topk(cos, 12, dim=1)
    assert torch.allclose(sim, rs, rtol=1e-4, atol=1e-5)
    assert (idx == ri).float().mean() > 0.95              # ties may permute
    # a large query batch goes through the tcgen05 (tf32) screen + exact fp32 re-rank
    q16 = torch.randn(16, d)
    idx16, sim16 = eng.top_k(q16, 10)
    qn16 = q16 / q16.norm(dim=1, keepdim=True)
    cos16 = (qn16 @ full.t()) / torch.where(nr > 0, nr, torch.ones_like(nr))
    cos16[:, nr == 0] = 0
    rs16, ri16 = torch.topk(cos16, 10, dim=1)
    assert torch.allclose(sim16, rs16, rtol=1e-4, atol=1e-5)
    assert (idx16 == ri16).float().mean() > 0.95


def test_topk_threshold_filter_is_exact(monkeypatch):
    """Candidate selection by sampled threshold + one filter pass must return exactly what the per-chunk
    arg-max path returns - including on data that overflows the candidate buffer (falls back)."""
    dev = _dev()
    C = _C()
    g = torch.Generator().manual_seed(3)
    v, q, k = 300000, 7, 10
    scores = torch.randn(q, v, generator=g).to(dev)
    norms = (torch.rand(v, generator=g) + 0.5).to(dev)
    norms[123] = 0.0
    cos = torch.where(norms > 0, scores / norms, torch.zeros_like(scores))
    want_v, want_i = torch.topk(cos, k, dim=1)
    monkeypatch.setenv("GW2V_TOPK_FILTER", "1")
    i1, v1 = C.cosine_topk(scores.clone(), norms, k)
    monkeypatch.setenv("GW2V_TOPK_FILTER", "0")
    i0, v0 = C.cosine_topk(scores.clone(), norms, k)
    assert torch.equal(v1, want_v) and torch.equal(v0, want_v)
    assert torch.equal(i1, want_i) and torch.equal(i0, want_i)
    # adversarial: every row ties -> every element passes the threshold -> buffer overflow -> exact fallback
    monkeypatch.setenv("GW2V_TOPK_FILTER", "1")
    flat = torch.ones(q, v, device=dev)
    it, vt = C.cosine_topk(flat, torch.ones(v, device=dev), k)
    assert torch.equal(vt, torch.ones(q, k, device=dev))
    assert int(it.min()) >= 0 and all(len(set(r.tolist())) == k for r in it.cpu())
    # rows sorted by score: the sample must still bound the true k-th best
    ramp = torch.arange(v, dtype=torch.float32, device=dev).repeat(2, 1)
    ir, vr = C.cosine_topk(ramp, torch.ones(v, device=dev), k)
    assert ir[0].tolist() == list(range(v - 1, v - 1 - k, -1))


def test_matrix_io_streams_from_and_to_the_device(tmp_path, monkeypatch):
    """save_matrix / load_matrix move the shard through a pinned double buffer in row chunks."""
    dev = _dev()
    from glint_word2vec_b200.models import matrix_io
    from glint_word2vec_b200.parallel.comm import Comm
    monkeypatch.setattr(matrix_io, "CHUNK_BYTES", 64 << 10)          # many chunks
    v, d = 5003, 100
    eng, _ = _make_engine(dev, v, d)
    eng.syn1 = (torch.randn(v, eng.shard.cols) * (torch.arange(eng.shard.cols) < d)).to(dev)
    matrix_io.save_matrix(eng, str(tmp_path / "m"))
    raw = np.load(str(tmp_path / "m" / "matrix" / "syn1.00of01.npy"))
    assert raw.shape == (v, d) and np.array_equal(raw, eng.syn1[:, :d].cpu().numpy())
    back = matrix_io.load_matrix(str(tmp_path / "m"), Comm(), dev)
    assert back.syn0.is_cuda and torch.equal(back.syn0, eng.syn0) and torch.equal(back.syn1, eng.syn1)


def test_fit_on_gpu_golden(corpus_sentences):
    """Scenario 1/10/12 of the reference spec on the device path (SPEC:83-106,290-352)."""
    _dev()
    from glint_word2vec_b200 import ServerSideGlintWord2Vec
    est = ServerSideGlintWord2Vec(seed=1, stepSize=0.025, numPartitions=2, numParameterServers=1,
                                  inputCol="sentence", outputCol="model", unigramTableSize=1000000,
                                  parameterServerConfig={"subsample_mode": "reference", "max_hot_updates": 16})
    model = est.fit({"sentence": corpus_sentences})
    try:
        syn = model.findSynonymsArray("österreich", 10)
        assert len(syn) == 10
        words = [w for w, _ in syn]
        assert "wien" in words
        assert dict(syn)["wien"] > 0.9
        v = model.transformWord("wien") - model.transformWord("österreich") + model.transformWord("deutschland")
        an = dict(model.findSynonymsArray(v, 10))
        assert "berlin" in an and an["berlin"] > 0.9
    finally:
        model.stop()


@pytest.mark.parametrize("k,q", [(64, 1), (64, 16), (128, 17), (512, 64), (256, 200), (32, 256)])
def test_scores_tc_tcgen05_matches_fp32(k, q):
    """tcgen05 (kind::tf32, TMEM accumulators, TMA operands) score GEMM vs an fp32 matmul."""
    dev = _dev()
    C = _C()
    assert C.scores_tc_supported(k, q)
    v = 30011                                            # not a multiple of the 128-row tile
    g = torch.Generator().manual_seed(k * 1000 + q)
    syn0 = torch.randn(v, k, generator=g).to(dev)
    qs = torch.randn(q, k, generator=g).to(dev)
    out = C.scores_tc(syn0, qs)
    torch.cuda.synchronize()
    ref = (qs.double() @ syn0.double().t()).float()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err / scale < 2e-3, (err, scale)              # tf32 operands (10-bit mantissa), fp32 accumulate
    # the exact CUDA-core path agrees to fp32 rounding
    out2 = C.scores_rows(syn0, qs[:8].contiguous())
    assert torch.allclose(out2, ref[:8], rtol=1e-4, atol=1e-4)


def test_pairs_kernel_hot_row_damping_matches_oracle_with_row_scales():
    """Hot-row damping in the pair kernel: the update of row r is scaled by the engine's table for the first H rows.
    Checked against the oracle given the same scales, on tokens that include the hot rows."""
    dev = _dev()
    v, d, t = 100000, 64, 3000
    cfg = SGNSConfig(v, d, 5, 5, seed=7)
    eng = ShardEngine(cfg, device=dev, options=EngineOptions(subsample_mode="reference", hot_row_cap=4.0))
    eng.init_weights()
    eng.set_noise(zipf_counts(v, 10 ** 7, 1.0))
    g = torch.Generator().manual_seed(0)
    syn1 = torch.randn(v, d, generator=g) * (0.5 / d ** 0.5)
    syn0 = torch.randn(v, d, generator=g) * (0.5 / d ** 0.5)
    eng.syn0, eng.syn1 = syn0.to(dev), syn1.to(dev)
    s0, s1 = eng.row_scales(eng.inflight_tokens(t))
    h = s0.shape[0]
    assert 10 < h < v and s0[0] < 0.05
    full0, full1 = torch.ones(v), torch.ones(v)
    full0[:h], full1[:h] = torch.from_numpy(s0), torch.from_numpy(s1)
    rng = np.random.default_rng(1)
    tokens = np.concatenate([np.arange(0, 400), rng.choice(np.arange(400, v), size=t - 400, replace=False)]).astype(np.int32)
    rng.shuffle(tokens)
    sid = (np.arange(t) // 37).astype(np.int32)
    ref0, ref1 = syn0.clone(), syn1.clone()
    st = sgns.sgns_minibatch_reference(ref0, ref1, cfg, eng.alias, tokens, sid, 5, 0, 0.002, row_scale0=full0, row_scale1=full1)
    plain0 = syn0.clone()
    sgns.sgns_minibatch_reference(plain0, syn1.clone(), cfg, eng.alias, tokens, sid, 5, 0, 0.002)
    stats = eng.train_step(tokens, sid, 5, 0, 0.002).cpu()
    assert int(stats[0]) == st.pairs
    got0, got1 = eng.syn0.cpu(), eng.syn1.cpu()
    r0, r1 = ref0 - syn0, ref1 - syn1
    assert float((got0 - syn0 - r0).norm() / r0.norm()) < 2e-2
    assert float((got1 - syn1 - r1).norm() / r1.norm()) < 2e-2
    assert float((plain0 - syn0 - r0).norm() / r0.norm()) > 0.1          # the scales do change the update


@pytest.mark.parametrize("d,nq,k", [(64, 1, 10), (64, 64, 10), (100, 17, 5), (300, 64, 40), (128, 200, 1)])
def test_nn_select_matches_exact_topk(d, nq, k):
    """Fused select path (csrc/nn_select.cu + ops/nn.py): tcgen05 screening with in-epilogue threshold selection,
    exact fp32 re-score -- against a dense fp32 cosine top-k.  d = 100 / 300: rows are zero-padded to 32 floats."""
    dev = _dev()
    v = 300000
    eng = ShardEngine(SGNSConfig(v, d, 5, 5, seed=3), device=dev, options=EngineOptions(hot_row_cap=0))
    eng.init_weights()
    g = torch.Generator().manual_seed(1)
    w = torch.randn(v, eng.shard.cols, generator=g)
    w[:, d:] = 0
    w[7] = 0                                   # a zero row scores 0, like the reference
    w[1000:1010] = w[5] * torch.linspace(0.5, 2.0, 10)[:, None]     # exact cosine ties with row 5
    eng.syn0 = w.to(dev)
    eng._touch()
    qs = torch.cat([w[:nq // 2 + 1, :d] + 0.3 * torch.randn(nq // 2 + 1, d, generator=g),
                    torch.randn(nq - nq // 2 - 1, d, generator=g)])[:nq]
    idx, sim = eng.top_k(qs, k)
    nn = eng._cuda.nn_index()
    assert nn.version == eng._version and nn.overflows == 0      # the select path answered
    wn = w[:, :d].to(dev)
    cos = (qs.to(dev) / qs.to(dev).norm(dim=1, keepdim=True)) @ (wn / wn.norm(dim=1, keepdim=True).clamp(min=1e-30)).t()
    want_sim, want_idx = torch.topk(cos, k, dim=1)
    assert torch.allclose(sim, want_sim.cpu(), atol=3e-6)
    # the returned rows really have those cosines (rows 1000..1009 tie exactly with row 5, so index SETS may differ)
    got_cos = torch.gather(cos.cpu(), 1, idx)
    assert torch.allclose(got_cos, want_sim.cpu(), atol=3e-6)
    assert all(len(set(r.tolist())) == k for r in idx)
    # training invalidates the index
    eng.set_noise(zipf_counts(v, 10 ** 7, 0.6))
    eng.train_step(np.arange(1000, dtype=np.int32), np.zeros(1000, np.int32), 0, 0, 0.025)
    eng.top_k(qs[:1], k)
    assert nn.version == eng._version


def test_nn_select_overflow_falls_back_to_dense():
    """All rows identical: every cosine passes the threshold, the candidate lists overflow, the dense path answers."""
    dev = _dev()
    v, d = 300000, 64
    eng = ShardEngine(SGNSConfig(v, d, 5, 5, seed=3), device=dev, options=EngineOptions(hot_row_cap=0))
    eng.init_weights()
    eng.syn0 = torch.ones(v, d, device=dev)
    eng._touch()
    idx, sim = eng.top_k(torch.ones(2, d), 5)
    assert eng._cuda.nn_index().overflows == 1
    assert torch.allclose(sim, torch.ones(2, 5), atol=1e-5)
