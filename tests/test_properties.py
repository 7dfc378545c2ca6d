"""Property-based tests (hypothesis) of the host-side building blocks (unit tier)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from glint_word2vec_b200.data.corpus import chunk_encoded, iter_steps, java_split
from glint_word2vec_b200.data.sampler import build_alias
from glint_word2vec_b200.models import sgns
from glint_word2vec_b200.models.sgns import SGNSConfig
from glint_word2vec_b200.parallel.sharding import make_shard, shard_cols
from glint_word2vec_b200.utils import philox


def _java_split_reference(line: str):
    """String.split(" ") of the JVM: split on every single space, then drop TRAILING empty strings;
    an input without any separator is returned as is (so "" -> [""])."""
    parts = line.split(" ")
    if len(parts) == 1:
        return parts
    while parts and parts[-1] == "":
        parts.pop()
    return parts


@given(st.text(alphabet=st.sampled_from(["a", "b", " ", "ö"]), max_size=30))
def test_java_split_matches_jvm_semantics(line):
    assert java_split(line) == _java_split_reference(line)


@given(st.lists(st.floats(min_value=1e-3, max_value=1e3, allow_nan=False), min_size=1, max_size=200),
       st.booleans())
@settings(max_examples=60, deadline=None)
def test_alias_table_reproduces_the_distribution(weights, native):
    w = np.asarray(weights, dtype=np.float64)
    t = build_alias(w, use_native=native)
    p = t.probabilities()
    assert p.shape == w.shape
    assert abs(p.sum() - 1.0) < 1e-6
    assert np.allclose(p, w / w.sum(), atol=2e-6)          # uint32 threshold quantisation: 2^-32 per bucket


@given(st.integers(0, 2 ** 63 - 1), st.integers(0, 7), st.integers(0, 2 ** 40), st.integers(0, 2 ** 31), st.integers(0, 1000))
@settings(max_examples=50, deadline=None)
def test_philox_vector_equals_scalar(seed, stream, pos, sub, iteration):
    """The vectorised generator is elementwise the scalar one, and distinct counters give distinct words."""
    posv = np.array([pos, pos + 1, pos], dtype=np.uint64)
    subv = np.array([sub, sub, sub + 1], dtype=np.uint64)
    a = philox.rand4(seed, stream, posv, subv, iteration)
    for j in range(3):
        b = philox.rand4(seed, stream, np.array([posv[j]], dtype=np.uint64), np.array([subv[j]], dtype=np.uint64),
                         iteration)
        assert all(int(a[c][j]) == int(b[c][0]) for c in range(4))
    assert tuple(int(a[c][0]) for c in range(4)) != tuple(int(a[c][1]) for c in range(4))
    assert tuple(int(a[c][0]) for c in range(4)) != tuple(int(a[c][2]) for c in range(4))


@given(st.lists(st.integers(0, 40), min_size=1, max_size=30), st.integers(1, 12))
def test_chunking_preserves_tokens_and_bounds_lengths(lengths, max_len):
    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    tokens = np.arange(offsets[-1], dtype=np.int32)
    c = chunk_encoded(tokens, offsets, max_len)
    assert np.array_equal(c.tokens, tokens)
    lens = np.diff(c.offsets)
    assert (lens <= max_len).all() and (lens > 0).all()              # no over-long and no empty sentences
    assert c.offsets[0] == 0 and c.offsets[-1] == offsets[-1]
    # chunk boundaries refine the sentence boundaries
    assert set(offsets.tolist()) <= set(c.offsets.tolist())


@given(st.lists(st.integers(1, 60), min_size=1, max_size=40), st.integers(8, 200))
def test_steps_cover_the_corpus_with_whole_sentences(lengths, step_tokens):
    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    tokens = np.arange(offsets[-1], dtype=np.int32)
    from glint_word2vec_b200.data.corpus import EncodedCorpus
    corpus = EncodedCorpus(tokens, offsets)
    seen = []
    pos = 0
    for b in iter_steps(corpus, step_tokens):
        assert b.raw_pos0 == pos
        assert len(b.tokens) == len(b.sent_id) == b.n_words
        assert 0 < b.n_words <= step_tokens
        if max(lengths) <= step_tokens:
            # a step holds whole sentences: its first and last token sit on sentence boundaries
            assert pos in set(offsets.tolist()) and (pos + b.n_words) in set(offsets.tolist())
        # sentence ids inside the step are monotone and change exactly at sentence boundaries of the corpus
        cuts = np.flatnonzero(np.diff(b.sent_id)) + 1 + pos
        assert set(cuts.tolist()) <= set(offsets.tolist())
        seen.append(b.tokens)
        pos += b.n_words
    assert np.array_equal(np.concatenate(seen), tokens)


@given(st.integers(1, 700), st.sampled_from([1, 2, 3, 4, 5, 8]))
def test_column_shards_tile_the_vector(d, world):
    k = shard_cols(d, world)
    assert k % 4 == 0 and k * world >= d
    covered = 0
    for r in range(world):
        sh = make_shard(d, world, r)
        assert sh.cols == k and sh.col_start == min(r * k, d)
        assert 0 <= sh.real_cols <= k
        covered += sh.real_cols
    assert covered == d


@given(st.integers(2, 9), st.integers(0, 3), st.integers(1, 300))
@settings(max_examples=40, deadline=None)
def test_pairs_are_symmetric_under_reference_window_bounds(window, iteration, t):
    """Every (centre, context) pair lies inside one sentence, inside the drawn window, never pairs a token
    with itself, and the slot encodes the offset."""
    cfg = SGNSConfig(1000, 8, window, 3, seed=5)
    rng = np.random.default_rng(t)
    tokens = rng.integers(0, 1000, size=t).astype(np.int32)
    sid = np.cumsum(rng.random(t) < 0.1).astype(np.int32)
    ci, cj, slot = sgns.enumerate_pairs(cfg, tokens, sid, 99, iteration)
    assert (ci != cj).all()
    assert (sid[ci] == sid[cj]).all()
    assert (np.abs(cj - ci) <= window).all()
    assert np.array_equal(slot, (cj - ci) + window)
