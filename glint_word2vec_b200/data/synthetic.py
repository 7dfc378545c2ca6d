"""Small synthetic corpora with planted structure (tests, smoke runs, examples).

The reference's integration corpus (German Wikipedia articles about six
countries and their capitals, SPEC:22-36) is not redistributable with this
repository; when it is not mounted the golden tests fall back to this
generator, which plants the same kind of structure.
"""
from __future__ import annotations


def synthetic_capitals_corpus(n_sent=6000, seed=0):
    """Planted structure: every (country, capital) pair shares 8 pair-specific context
    words, so the two end up as nearest neighbours; analogies hold because all
    countries share one marker word and all capitals another."""
    import random
    rng = random.Random(seed)
    pairs = [("österreich", "wien"), ("deutschland", "berlin"), ("frankreich", "paris"),
             ("spanien", "madrid"), ("finnland", "helsinki"), ("grossbritannien", "london")]
    filler = [f"w{i}" for i in range(300)]
    sents = []
    for _ in range(n_sent):
        pi = rng.randrange(len(pairs))
        c, k = pairs[pi]
        local = [f"ctx{pi}_{j}" for j in range(8)]
        s = [rng.choice(filler) for _ in range(rng.randint(2, 8))]
        core = rng.sample(local, 3)
        if rng.random() < 0.8:
            core.append(c)
            core.append("land")
        if rng.random() < 0.8:
            core.append(k)
            core.append("stadt")
        rng.shuffle(core)
        pos = rng.randint(0, len(s))
        s[pos:pos] = core
        sents.append(s)
    return sents


def planted_pairs_corpus(vocab_size: int, n_tokens: int, n_pairs: int = 200, seed: int = 0, sent_len: int = 20,
                         plant_prob: float = 0.35, zipf: float = 1.0):
    """Index-encoded Zipf corpus with planted synonym pairs (quality gate of the device kernels).

    Background: Zipf(``zipf``) tokens over ``vocab_size`` words in sentences of ``sent_len``.  Planted structure:
    ``n_pairs`` pairs (a_i, b_i) of mid-frequency words, each with four private context words; with probability
    ``plant_prob`` a sentence receives a block ``ctx ctx X ctx ctx`` where X is a_i or b_i -- the two words share
    their contexts and nothing else does, so a trained model puts b_i among the nearest neighbours of a_i.

    Returns ``(tokens int32, offsets int64, counts int64, pairs int64 [n_pairs, 2])``."""
    import numpy as np
    rng = np.random.default_rng(seed)
    ranks = np.arange(1, vocab_size + 1, dtype=np.float64)
    p = ranks ** (-zipf)
    p /= p.sum()
    n_sent = max(1, n_tokens // sent_len)
    base = rng.choice(vocab_size, size=(n_sent, sent_len), p=p).astype(np.int32)
    # planted words live in a band of the vocabulary that the background rarely produces
    lo = min(vocab_size // 4, 2000)
    need = 6 * n_pairs
    if lo + need > vocab_size:
        raise ValueError("vocabulary too small for the planted structure")
    ids = lo + rng.permutation(need)
    pairs = np.stack([ids[:n_pairs], ids[n_pairs:2 * n_pairs]], 1).astype(np.int64)
    ctx = ids[2 * n_pairs:].reshape(n_pairs, 4)
    planted = rng.random(n_sent) < plant_prob
    which = rng.integers(0, n_pairs, size=n_sent)
    side = rng.integers(0, 2, size=n_sent)
    pos = rng.integers(0, sent_len - 5, size=n_sent)
    rows = np.nonzero(planted)[0]
    for k, j in enumerate((0, 1, None, 2, 3)):
        col = pos[rows] + k
        base[rows, col] = pairs[which[rows], side[rows]] if j is None else ctx[which[rows], j]
    tokens = base.reshape(-1)
    offsets = np.arange(0, n_sent * sent_len + 1, sent_len, dtype=np.int64)
    counts = np.maximum(np.bincount(tokens, minlength=vocab_size), 1).astype(np.int64)
    return tokens, offsets, counts, pairs


def planted_recall(vectors, pairs, k: int = 10) -> float:
    """Fraction of planted pairs (a, b) with b among the k nearest neighbours (cosine) of a."""
    import numpy as np
    v = np.asarray(vectors, dtype=np.float32)
    n = v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-30)
    hits = 0
    for a, b in np.asarray(pairs):
        sims = n @ n[a]
        sims[a] = -2.0
        top = np.argpartition(-sims, k)[:k]
        hits += int(b in top)
    return hits / max(1, len(pairs))
