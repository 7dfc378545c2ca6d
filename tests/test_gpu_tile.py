"""GPU numerics of the tensor-core training kernel (csrc/sgns_tile.cu, neg_sharing="tile") against the fp32 oracle
``models/sgns.py::_minibatch_tile_reference`` and of its descriptor assumptions (csrc/umma_probe.cu)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from glint_word2vec_b200.data.sampler import zipf_counts
from glint_word2vec_b200.models import sgns
from glint_word2vec_b200.models.engine import EngineOptions, ShardEngine
from glint_word2vec_b200.models.sgns import SGNSConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def test_umma_descriptor_probes():
    """Every shared-memory layout / descriptor of sgns_tile.cu, checked against numpy with exact integer data:
    K-major SWIZZLE_128B, MN-major SWIZZLE_128B_BASE32B (A and B), the dU / dV instruction sequences and TMA
    gather4 in both swizzle modes.  Each case runs in its own process (a bad descriptor kills the context)."""
    _dev()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks", "probe_umma3.py")], capture_output=True, text=True,
                       timeout=900)
    res = dict(l.split()[1:3] for l in r.stdout.splitlines() if l.startswith("RESULT") and len(l.split()) >= 3)
    for need in ("gather4_sw32", "b_mn_l1024_s512", "b_mn64_st4096_l4096_s512", "a_mn_st16384_l16384_s512", "dU_seq_s512",
                 "dV_seq"):
        assert res.get(need) == "PASS", (need, r.stdout[-2000:], r.stderr[-2000:])


def _engine(dev, v, d, nn, window=5, n=5, wmode="reference", seed=7):
    cfg = SGNSConfig(v, d, window, n, seed=seed, window_mode=wmode, neg_sharing="tile", tile_centres=128, tile_negatives=nn)
    eng = ShardEngine(cfg, device=dev, options=EngineOptions(subsample_mode="reference", hot_row_cap=0))
    eng.init_weights()
    counts = zipf_counts(v, 10 ** 7, 0.6)
    eng.set_noise(counts)
    g = torch.Generator().manual_seed(0)
    syn1 = torch.randn(v, eng.shard.cols, generator=g) * (0.5 / d ** 0.5)
    syn0 = torch.randn(v, eng.shard.cols, generator=g) * (0.5 / d ** 0.5)
    syn0[:, d:] = 0
    syn1[:, d:] = 0
    eng.syn0, eng.syn1 = syn0.to(dev), syn1.to(dev)
    return eng, syn0, syn1


def _check(eng, syn0, syn1, tokens, sid, alpha, pos0=12345, it=1, tol=1e-2):
    d = eng.cfg.vector_size
    ref0, ref1 = syn0[:, :d].clone(), syn1[:, :d].clone()
    st = sgns.sgns_minibatch_reference(ref0, ref1, eng.cfg, eng.alias, tokens, sid, pos0, it, alpha)
    stats = eng.train_step(tokens, sid, pos0, it, alpha).cpu()
    torch.cuda.synchronize()
    assert int(stats[0]) == st.pairs
    assert int(stats[3]) == len(tokens)
    assert abs(float(stats[1]) - st.loss) / max(st.loss, 1e-9) < 5e-3, (float(stats[1]), st.loss)
    got0, got1 = eng.syn0.cpu()[:, :d], eng.syn1.cpu()[:, :d]
    d0, d1 = got0 - syn0[:, :d], got1 - syn1[:, :d]
    r0, r1 = ref0 - syn0[:, :d], ref1 - syn1[:, :d]
    assert r0.abs().sum() > 0 and r1.abs().sum() > 0
    e0 = float((d0 - r0).norm() / r0.norm())
    e1 = float((d1 - r1).norm() / r1.norm())
    assert e0 < tol and e1 < tol, (e0, e1)
    # rows the oracle does not touch are not touched by the kernel either
    assert float(d0[r0.abs().sum(1) == 0].abs().sum()) == 0.0
    assert float(d1[r1.abs().sum(1) == 0].abs().sum()) == 0.0
    if eng.shard.cols > d:
        assert float(eng.syn0[:, d:].abs().sum()) == 0.0 and float(eng.syn1[:, d:].abs().sum()) == 0.0
    return st


@pytest.mark.parametrize("d,nn,window,n,wmode", [(64, 32, 5, 5, "reference"), (64, 64, 5, 5, "reference"),
                                                 (40, 32, 5, 5, "reference"), (128, 64, 3, 7, "word2vec_c"),
                                                 (300, 32, 5, 10, "word2vec_c"), (512, 64, 5, 5, "reference")])
def test_tile_kernel_one_tile_with_duplicates(d, nn, window, n, wmode):
    """One tile (<= 128 centres) drawn from a tiny vocabulary: the same word appears many times as centre, context and
    negative.  Inside a tile the kernel has the reference's exact mini-batch semantics (all dots from pre-update rows,
    summed updates), so it must equal the oracle up to tf32 rounding whatever the duplication."""
    dev = _dev()
    eng, syn0, syn1 = _engine(dev, 50, d, nn, window, n, wmode)
    rng = np.random.default_rng(3)
    for t in (128, 77, 1):
        tokens = rng.integers(0, 50, size=t).astype(np.int32)
        sid = (np.arange(t) // 23).astype(np.int32)
        eng.syn0, eng.syn1 = syn0.to(dev), syn1.to(dev)
        if t == 1:                                   # a single token has no context: a clean no-op (Q4)
            stats = eng.train_step(tokens, sid, 5, 0, 0.025).cpu()
            assert int(stats[0]) == 0 and torch.equal(eng.syn0.cpu(), syn0) and torch.equal(eng.syn1.cpu(), syn1)
        else:
            _check(eng, syn0, syn1, tokens, sid, 0.025)


def test_tile_kernel_dot_products_match_fp32():
    """The S tile itself (band window + shared negatives of tile 0) against fp32 dots: tf32 rounding only."""
    dev = _dev()
    d, nn, v = 128, 32, 5000
    eng, syn0, syn1 = _engine(dev, v, d, nn)
    rng = np.random.default_rng(5)
    t = 300
    tokens = rng.integers(0, v, size=t).astype(np.int32)
    sid = (np.arange(t) // 40).astype(np.int32)
    ops = eng._cuda
    ops.tile_dbg = torch.zeros(128 * (160 + nn), device=dev)
    eng.train_step(tokens, sid, 999, 2, 0.0)                       # alpha 0: weights unchanged
    S = ops.tile_dbg.cpu().view(128, 160 + nn)
    ops.tile_dbg = None
    tneg = sgns.tile_negatives(eng.cfg, eng.alias, 999, np.array([0]), 2)[0]
    U = syn0[tokens[:128].astype(np.int64), :d]
    want_neg = U @ syn1[tneg.astype(np.int64), :d].T
    assert float((S[:, 160:] - want_neg).abs().max()) < 2e-3 * float(want_neg.abs().max() + 1)
    for i in (0, 5, 64, 127):
        for off in (-2, -1, 1, 2):
            j = i + off
            if 0 <= j < t:
                want = float(U[i] @ syn1[int(tokens[j]), :d])
                assert abs(float(S[i, i + 16 + off]) - want) < 2e-3 * (abs(want) + 1)


@pytest.mark.parametrize("d,nn,t", [(64, 32, 3000), (512, 64, 1500), (100, 64, 2000)])
def test_tile_kernel_many_tiles_match_oracle(d, nn, t):
    """Many tiles over distinct tokens of a large vocabulary: tiles almost never touch each other's rows, so the
    asynchronous kernel must match the summed whole-step oracle closely."""
    dev = _dev()
    v = 300000
    eng, syn0, syn1 = _engine(dev, v, d, nn)
    rng = np.random.default_rng(1)
    tokens = rng.choice(v, size=t, replace=False).astype(np.int32)
    sid = (np.arange(t) // 37).astype(np.int32)
    _check(eng, syn0, syn1, tokens, sid, 0.002, tol=2e-2)


def test_tile_kernel_sentence_boundaries_and_steps():
    """Sentence ends inside tiles and at tile edges, several consecutive steps (ring / barrier phases carry over
    inside a launch, workspaces are reused across launches)."""
    dev = _dev()
    v, d, nn = 100000, 64, 32
    eng, syn0, syn1 = _engine(dev, v, d, nn)
    rng = np.random.default_rng(9)
    ref0, ref1 = syn0[:, :d].clone(), syn1[:, :d].clone()
    pos = 0
    for step, t in enumerate((128 * 3, 1000, 129, 4096)):
        tokens = rng.choice(v, size=t, replace=False).astype(np.int32)
        lens = rng.integers(1, 60, size=t)
        sid = np.repeat(np.arange(t), lens)[:t].astype(np.int32)
        sid[127:129] = sid[127] if step == 0 else sid[127:129]
        st = sgns.sgns_minibatch_reference(ref0, ref1, eng.cfg, eng.alias, tokens, sid, pos, 0, 0.002)
        stats = eng.train_step(tokens, sid, pos, 0, 0.002).cpu()
        assert int(stats[0]) == st.pairs
        pos += t
    got0, got1 = eng.syn0.cpu()[:, :d], eng.syn1.cpu()[:, :d]
    r0, r1 = ref0 - syn0[:, :d], ref1 - syn1[:, :d]
    assert float((got0 - syn0[:, :d] - r0).norm() / r0.norm()) < 3e-2
    assert float((got1 - syn1[:, :d] - r1).norm() / r1.norm()) < 3e-2


@pytest.mark.parametrize("world,d,nn,wmode,window", [(2, 64, 32, "reference", 5), (8, 128, 64, "reference", 5),
                                                     (4, 100, 32, "word2vec_c", 7)])
def test_tile_kernel_exchange_protocol_loopback(world, d, nn, wmode, window, monkeypatch):
    """The column-shard protocol of the tile kernel (payload extraction from TMEM, slots, release/acquire flags, ordered
    sum, one-tile lead of the push) on ONE GPU: with GW2V_LOOPBACK_WORLD = S every push lands in the GPU's own
    exchange buffer as the message of "rank" r carrying 1/S of the dots, so the result must equal the single-shard
    oracle.  Several steps: the slot ring and the sequence numbers carry over from launch to launch."""
    dev = _dev()
    monkeypatch.setenv("GW2V_LOOPBACK_WORLD", str(world))
    v = 200000
    eng, syn0, syn1 = _engine(dev, v, d, nn, window=window, wmode=wmode)
    rng = np.random.default_rng(4)
    ref0, ref1 = syn0[:, :d].clone(), syn1[:, :d].clone()
    pos = 0
    for t in (700, 128 * 148 * 2 + 5, 100, 3000):
        tokens = rng.choice(v, size=t, replace=False).astype(np.int32)
        sid = (np.arange(t) // 29).astype(np.int32)
        st = sgns.sgns_minibatch_reference(ref0, ref1, eng.cfg, eng.alias, tokens, sid, pos, 0, 0.002)
        stats = eng.train_step(tokens, sid, pos, 0, 0.002).cpu()
        assert int(stats[0]) == st.pairs
        assert abs(float(stats[1]) - st.loss) / st.loss < 5e-3
        pos += t
    got0, got1 = eng.syn0.cpu()[:, :d], eng.syn1.cpu()[:, :d]
    r0, r1 = ref0 - syn0[:, :d], ref1 - syn1[:, :d]
    assert float((got0 - syn0[:, :d] - r0).norm() / r0.norm()) < 3e-2
    assert float((got1 - syn1[:, :d] - r1).norm() / r1.norm()) < 3e-2
    assert int(eng._cuda._xchg["err"].item()) == 0


@pytest.mark.parametrize("d,nn,grid", [(128, 64, 4), (128, 32, 3), (300, 64, 2), (64, 32, 1)])
def test_tile_kernel_many_tiles_per_cta(d, nn, grid, monkeypatch):
    """Few CTAs, many tiles each (what a 131 072-token step does on 148 SMs): the stage ring interleaves pass B of tile
    t with pass A of tile t+1, and with the 3-stage ring of NN = 64 a stage alternates between the two kinds.  A parity
    wait that skipped a phase once let an epilogue group read stale U rows here (fixed: per-accumulator `b_full`)."""
    dev = _dev()
    monkeypatch.setenv("GW2V_TILE_GRID", str(grid))
    v = 200000
    eng, syn0, syn1 = _engine(dev, v, d, nn)
    rng = np.random.default_rng(8)
    ref0, ref1 = syn0[:, :d].clone(), syn1[:, :d].clone()
    pos = 0
    for t in (3000, 1500):
        tokens = rng.choice(v, size=t, replace=False).astype(np.int32)
        sid = (np.arange(t) // 33).astype(np.int32)
        st = sgns.sgns_minibatch_reference(ref0, ref1, eng.cfg, eng.alias, tokens, sid, pos, 0, 0.002)
        stats = eng.train_step(tokens, sid, pos, 0, 0.002).cpu()
        assert int(stats[0]) == st.pairs
        pos += t
    got0, got1 = eng.syn0.cpu()[:, :d], eng.syn1.cpu()[:, :d]
    r0, r1 = ref0 - syn0[:, :d], ref1 - syn1[:, :d]
    assert float((got0 - syn0[:, :d] - r0).norm() / r0.norm()) < 2e-2
    assert float((got1 - syn1[:, :d] - r1).norm() / r1.norm()) < 2e-2


def test_tile_kernel_hot_row_damping_matches_oracle_with_row_scales():
    dev = _dev()
    v, d, nn, t = 100000, 64, 32, 3000
    cfg = SGNSConfig(v, d, 5, 5, seed=7, neg_sharing="tile", tile_negatives=nn)
    eng = ShardEngine(cfg, device=dev, options=EngineOptions(subsample_mode="reference", hot_row_cap=4.0))
    eng.init_weights()
    eng.set_noise(zipf_counts(v, 10 ** 7, 1.0))
    g = torch.Generator().manual_seed(0)
    syn1 = torch.randn(v, d, generator=g) * (0.5 / d ** 0.5)
    syn0 = torch.randn(v, d, generator=g) * (0.5 / d ** 0.5)
    eng.syn0, eng.syn1 = syn0.to(dev), syn1.to(dev)
    s0, s1 = eng.row_scales(eng.inflight_tokens(t))
    h = s0.shape[0]
    full0, full1 = torch.ones(v), torch.ones(v)
    full0[:h], full1[:h] = torch.from_numpy(s0), torch.from_numpy(s1)
    rng = np.random.default_rng(1)
    tokens = np.concatenate([np.arange(0, 400), rng.choice(np.arange(400, v), size=t - 400, replace=False)]).astype(np.int32)
    rng.shuffle(tokens)
    sid = (np.arange(t) // 37).astype(np.int32)
    ref0, ref1 = syn0.clone(), syn1.clone()
    assert 0.0 < eng.tile_neg_scale() < 0.1            # 128 centres x 4 pairs x 5 / 32 = 80 unit updates per event, cap 4
    st = sgns.sgns_minibatch_reference(ref0, ref1, cfg, eng.alias, tokens, sid, 5, 0, 0.002, row_scale0=full0, row_scale1=full1,
                                       tile_neg_scale=eng.tile_neg_scale())
    stats = eng.train_step(tokens, sid, 5, 0, 0.002).cpu()
    assert int(stats[0]) == st.pairs
    got0, got1 = eng.syn0.cpu(), eng.syn1.cpu()
    r0, r1 = ref0 - syn0, ref1 - syn1
    assert float((got0 - syn0 - r0).norm() / r0.norm()) < 2e-2
    assert float((got1 - syn1 - r1).norm() / r1.norm()) < 2e-2
