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
