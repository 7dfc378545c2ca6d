"""Unit tier (CPU): Philox, vocabulary, tokenisation quirks, alias table,
sub-sampling thresholds, window generator, LR schedule."""
from collections import Counter

import numpy as np
import pytest

from glint_word2vec_b200.data import corpus as C
from glint_word2vec_b200.data import sampler as S
from glint_word2vec_b200.data.vocab import build_vocab, vocab_from_counts
from glint_word2vec_b200.models import sgns
from glint_word2vec_b200.utils import philox


def test_philox_known_answer():
    # Random123 known-answer test vectors for philox4x32-10
    r = philox.philox4x32(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    r = philox.philox4x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(x) for x in r] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    r = philox.philox4x32(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [int(x) for x in r] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_streams_differ_and_reproduce():
    a = philox.rand4(7, philox.STREAM_NEG, np.arange(100), 3, iteration=1)
    b = philox.rand4(7, philox.STREAM_NEG, np.arange(100), 3, iteration=1)
    c = philox.rand4(7, philox.STREAM_WINDOW, np.arange(100), 3, iteration=1)
    d = philox.rand4(7, philox.STREAM_NEG, np.arange(100), 3, iteration=2)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert not np.array_equal(a[0], c[0]) and not np.array_equal(a[0], d[0])


def test_java_split_quirks():
    assert C.java_split("a b") == ["a", "b"]
    assert C.java_split("a  b") == ["a", "", "b"]          # interior empty token kept (Q9)
    assert C.java_split(" a") == ["", "a"]
    assert C.java_split("a b  ") == ["a", "b"]             # trailing empties dropped
    assert C.java_split("") == [""]                         # no match -> the input itself
    assert C.java_split(" ") == []


@pytest.mark.parametrize("native", [False, True])
def test_vocab_matches_counter(native):
    sents = [["b", "a", "a", "", "c"], ["a", "b", "", "zz"], ["c", "c", "c"], []]
    v = build_vocab(sents, min_count=2, use_native=native)
    cnt = Counter(w for s in sents for w in s)
    expect = sorted([(w, c) for w, c in cnt.items() if c >= 2], key=lambda wc: (-wc[1], wc[0].encode()))
    assert v.words == [w for w, _ in expect]
    assert v.counts.tolist() == [c for _, c in expect]
    assert "" in v.index                                    # the empty string is a legal word (Q9)
    assert v.train_words == sum(c for _, c in expect)
    with pytest.raises(ValueError):
        build_vocab(sents, min_count=100, use_native=native)


def test_reference_corpus_vocab(corpus_sentences):
    import os
    from conftest import CORPUS
    if not os.path.exists(CORPUS):
        pytest.skip("reference corpus not mounted")
    v = build_vocab(corpus_sentences, 5)
    assert v.size == 3611                                   # SPEC:33
    assert v.train_words == 118755
    assert v.counts[v.index[""]] == 2155


@pytest.mark.parametrize("native", [False, True])
def test_encode_chunks_and_drops_oov(native):
    sents = [["a", "x", "b", "a", "b", "a"], ["y"], ["b"]]
    v = build_vocab(sents, min_count=2, use_native=False)
    enc = C.encode_corpus(sents, v, max_sentence_length=2, use_native=native)
    ia, ib = v.index["a"], v.index["b"]
    assert enc.tokens.tolist() == [ia, ib, ia, ib, ia, ib]
    assert enc.offsets.tolist() == [0, 2, 4, 5, 6]


def test_iter_steps_packs_whole_sentences():
    toks = np.arange(25, dtype=np.int32)
    offs = np.array([0, 4, 9, 10, 18, 25], dtype=np.int64)
    steps = list(C.iter_steps(C.EncodedCorpus(toks, offs), 10))
    assert [s.tokens.tolist() for s in steps] == [list(range(0, 10)), list(range(10, 18)), list(range(18, 25))]
    assert steps[0].sent_id.tolist() == [0] * 4 + [1] * 5 + [2]
    assert [s.raw_pos0 for s in steps] == [0, 10, 18]
    big = list(C.iter_steps(C.EncodedCorpus(toks, np.array([0, 25])), 10))
    assert sum(s.n_words for s in big) == 25 and max(len(s.tokens) for s in big) <= 10


@pytest.mark.parametrize("native", [False, True])
def test_alias_table_distribution(native):
    rng = np.random.default_rng(0)
    counts = rng.integers(1, 1000, size=500)
    at = S.unigram_alias(counts, 0.75, use_native=native)
    p = counts.astype(np.float64) ** 0.75
    p /= p.sum()
    assert np.abs(at.probabilities() - p).max() < 1e-9
    n = 400000
    r = philox.rand4(5, philox.STREAM_ZIPF, np.arange(n))
    draws = at.sample(r[0], r[1])
    obs = np.bincount(draws, minlength=500)
    chi2 = ((obs - n * p) ** 2 / (n * p)).sum()
    assert chi2 < 500 + 6 * np.sqrt(2 * 500)               # chi-square, 499 dof


def test_keep_thresholds():
    counts = np.array([1000000, 1000, 10, 1])
    t = S.keep_thresholds(counts, 1e-3, "word2vec")
    f = counts / counts.sum()
    keep = np.minimum(1.0, (np.sqrt(f / 1e-3) + 1) * 1e-3 / f)
    assert np.allclose(t.astype(np.float64) / 2 ** 32, keep, atol=1e-6)
    assert (S.keep_thresholds(counts, 1e-3, "reference") == 0xFFFFFFFF).all()      # Q1: inert
    tok = np.zeros(200000, dtype=np.int32)
    m = sgns.subsample_mask(tok, t, seed=1, iteration=0, raw_pos0=0)
    assert abs(m.mean() - keep[0]) < 0.01


def test_window_modes_distribution():
    cfg = sgns.SGNSConfig(100, window=5, window_mode="reference")
    lo, hi = sgns.window_bounds(cfg, np.arange(200000, dtype=np.uint64), 0)
    b = -lo
    assert set(np.unique(b)) == {0, 1, 2, 3, 4} and np.array_equal(hi, b - 1)
    ncontexts = np.where(b > 0, 2 * b - 1, 0)
    assert abs(ncontexts.mean() - 3.2) < 0.05              # Q2: E[pairs/centre] = 3.2
    cfg2 = sgns.SGNSConfig(100, window=5, window_mode="word2vec_c")
    lo2, hi2 = sgns.window_bounds(cfg2, np.arange(200000, dtype=np.uint64), 0)
    assert set(np.unique(hi2)) == {1, 2, 3, 4, 5} and np.array_equal(lo2, -hi2)


def test_pairs_respect_sentences_and_window():
    cfg = sgns.SGNSConfig(50, window=3)
    toks = np.arange(40, dtype=np.int32) % 50
    sid = (np.arange(40) // 8).astype(np.int32)
    ci, cj, slot = sgns.enumerate_pairs(cfg, toks, sid, 1000, 0)
    assert (sid[ci] == sid[cj]).all() and (ci != cj).all()
    assert (np.abs(ci - cj) <= 2).all()                     # reference radius <= window-1
    assert np.array_equal(slot, cj - ci + 3)
    # a partial range enumerates exactly the same pairs for those centres
    ci2, cj2, _ = sgns.enumerate_pairs(cfg, toks, sid, 1000, 0, 8, 24)
    sel = (ci >= 8) & (ci < 24)
    assert np.array_equal(ci[sel], ci2) and np.array_equal(cj[sel], cj2)


def test_learning_rate_schedule():
    assert sgns.learning_rate(0.025, 0, 1000) == 0.025
    assert abs(sgns.learning_rate(0.025, 500, 1000) - 0.025 * (1 - 500 / 1001)) < 1e-12
    assert sgns.learning_rate(0.025, 10 ** 9, 1000) == 0.025 * 1e-4    # floor (MLLIB:410)


def test_sigmoid_table_mode_close_to_exact():
    import torch
    f = torch.linspace(-8, 8, 1001)
    a = sgns.sigmoid_coeff(f, 1.0, 1.0, "exact")
    b = sgns.sigmoid_coeff(f, 1.0, 1.0, "table")
    assert (a - b).abs().max() < 0.012        # table step + the 83.0 index-scale quirk (MLLIB:300)
    assert float(sgns.sigmoid_coeff(torch.tensor([7.0]), 1.0, 1.0)) == 0.0
    assert float(sgns.sigmoid_coeff(torch.tensor([-7.0]), 0.0, 1.0)) == 0.0


def test_synthetic_vocab_is_lazy():
    v = vocab_from_counts(np.arange(10, 0, -1))
    assert v.size == 10 and v.words[3] == "w3" and v.index["w7"] == 7 and "w11" not in v.index


def test_neg_sharing_centre_shares_negatives_per_centre():
    """neg_sharing="centre": one draw of n negatives per centre position, reused by all of its pairs;
    "pair" (reference behaviour): private negatives per pair."""
    from glint_word2vec_b200.data.sampler import build_alias, zipf_counts
    from glint_word2vec_b200.models.sgns import SGNSConfig
    v = 5000
    alias = build_alias(zipf_counts(v, 10 ** 6).astype(np.float64))
    rng = np.random.default_rng(0)
    tokens = rng.integers(0, v, size=400).astype(np.int32)
    sid = (np.arange(400) // 50).astype(np.int32)
    for mode in ("pair", "centre"):
        cfg = SGNSConfig(v, 16, 5, 5, seed=9, neg_sharing=mode)
        ci, cj, slot = sgns.enumerate_pairs(cfg, tokens, sid, 777, 0)
        negs = sgns.draw_negatives(cfg, alias, np.uint64(777) + ci.astype(np.uint64), slot, 0)
        same = 0
        total = 0
        for a in range(1, len(ci)):
            if ci[a] == ci[a - 1]:
                total += 1
                same += int(np.array_equal(negs[a], negs[a - 1]))
        assert total > 100
        assert same == (total if mode == "centre" else 0) or (mode == "pair" and same < 3)
    with pytest.raises(ValueError):
        SGNSConfig(v, 16, neg_sharing="batch")


@pytest.mark.parametrize("tokenizer", ["java", "whitespace"])
def test_text_file_loader_matches_in_memory_path(tmp_path, tokenizer):
    """Native mmap/multi-thread file loader == pure-Python file reader == in-memory sentence path,
    including empty lines, leading/trailing/double spaces, CRLF and a missing final newline."""
    from glint_word2vec_b200.data.corpus import encode_corpus, encode_text_file, iter_text_file
    from glint_word2vec_b200.data.vocab import build_vocab, build_vocab_from_file
    from glint_word2vec_b200.ops import host
    rng = np.random.default_rng(0)
    words = [f"w{i}" for i in range(300)] + ["ö", "日本"]
    lines = []
    for i in range(60000):
        n = int(rng.integers(0, 12))
        toks = [words[int(rng.zipf(1.5)) % len(words)] for _ in range(n)]
        sep = "  " if i % 97 == 0 else " "
        line = sep.join(toks)
        if i % 53 == 0:
            line = " " + line
        if i % 59 == 0:
            line = line + "  "
        lines.append(line)
    lines[10] = ""
    lines[11] = "   "
    body = "\n".join(lines[:30000]) + "\r\n" + "\n".join(lines[30000:])        # > 1 MiB -> several threads
    path = tmp_path / "corpus.txt"
    path.write_bytes(body.encode("utf-8"))
    sentences = list(iter_text_file(str(path), tokenizer))
    assert len(sentences) == 60000
    ref_vocab = build_vocab(sentences, 3, use_native=False)
    ref_corpus = encode_corpus(sentences, ref_vocab, 7, use_native=False)
    for native in ([True, False] if host.available() else [False]):
        v = build_vocab_from_file(str(path), 3, tokenizer, use_native=native)
        assert v.words == ref_vocab.words and np.array_equal(v.counts, ref_vocab.counts)
        c = encode_text_file(str(path), v, 7, tokenizer, use_native=native)
        assert np.array_equal(c.tokens, ref_corpus.tokens)
        assert np.array_equal(c.offsets, ref_corpus.offsets)
    if tokenizer == "java":
        assert "" in ref_vocab.index                       # Q9: the empty token is a word
    else:
        assert "" not in ref_vocab.index
