import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CORPUS = "/root/reference/de_wikipedia_articles_country_capitals.txt"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session")
def corpus_sentences():
    """The reference's integration-test corpus (SPEC:22-36) when mounted, else a
    synthetic corpus with planted country/capital structure."""
    from glint_word2vec_b200.data.corpus import java_split
    if os.path.exists(CORPUS):
        with open(CORPUS, encoding="utf-8") as f:
            lines = f.read().split("\n")
        return [java_split(l) for l in lines]
    return synthetic_capitals_corpus()


from glint_word2vec_b200.data.synthetic import synthetic_capitals_corpus  # noqa: E402,F401


# property tests must not be flaky in CI: derandomised example generation, no deadline, no on-disk database
try:
    from hypothesis import settings as _hyp_settings
    _hyp_settings.register_profile("ci", derandomize=True, deadline=None, database=None)
    _hyp_settings.load_profile("ci")
except ImportError:                                   # hypothesis is optional
    pass


# denormal arithmetic is 50-100x slower on x86 and whether a torch worker thread flushes them depends on which thread
# loaded which library first; the CPU suite must not depend on that
try:
    import torch as _torch
    _torch.set_flush_denormal(True)
except Exception:  # pragma: no cover
    pass
