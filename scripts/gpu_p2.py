"""Where does the fit loop lose time?  Variants of the step loop on one GPU, wall ms/step (sync at the end)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from glint_word2vec_b200 import _C
from glint_word2vec_b200.data.corpus import chunk_encoded, iter_steps
from glint_word2vec_b200.data.sampler import build_alias, zipf_counts
from glint_word2vec_b200.models.engine import EngineOptions, ShardEngine
from glint_word2vec_b200.models.sgns import SGNSConfig
from glint_word2vec_b200.models import trainer as T
V, D, B, NS = 10_000_000, 512, 131072, 60
counts = zipf_counts(V, 200 * B, 1.0)
alias = build_alias(counts.astype(np.float64))
dev = torch.device("cuda", 0)
tok_dev = torch.empty(NS * B, dtype=torch.int32, device=dev)
_C.zipf_stream(torch.from_numpy(alias.packed()).to(dev), 2024, 0, tok_dev)
toks = tok_dev.cpu().numpy(); del tok_dev
corpus = chunk_encoded(toks, np.arange(0, NS * B + 1, 1000, dtype=np.int64), 1000)
eng = ShardEngine(SGNSConfig(V, D, 5, 5, seed=1), device=dev, options=EngineOptions(subsample_mode="word2vec", subsample_ratio=1e-4, step_tokens=B))
eng.init_weights(); eng.set_noise(counts)
eng._cuda.prepare(B); torch.cuda.synchronize()
def loop(name, steps_iter, use_async=True, n=None):
    torch.cuda.synchronize(); t0 = time.time(); k = 0; pend = []
    for b in steps_iter:
        h = eng.train_step_async(b.tokens, b.sent_id, b.raw_pos0, 0, 0.02) if use_async else eng.train_step(b.tokens, b.sent_id, b.raw_pos0, 0, 0.02)
        pend.append(h); k += 1
        if len(pend) > 2:
            x = pend.pop(0)
            if use_async: x.result()
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"{name:55s} {k} steps  {dt / k * 1e3:.3f} ms/step", flush=True)
for rep in range(2):
    loop("plain iter_steps (numpy -> stage copies)", iter_steps(corpus, B))
    loop("iter_steps + _pinned (main thread)", T._pinned(iter_steps(corpus, B), T._PinnedPool(B)))
    loop("_prefetch(iter_steps)", T._prefetch(iter_steps(corpus, B), 4))
    loop("_prefetch(_pinned(iter_steps))  [trainer]", T._prefetch(T._pinned(iter_steps(corpus, B), T._PinnedPool(B)), 4))
    b0 = next(iter(T._pinned(iter_steps(corpus, B), T._PinnedPool(B))))
    print("pinned?", b0.tokens.is_pinned(), b0.sent_id.is_pinned(), type(b0.tokens))
    class One:
        def __iter__(self): return (b0 for _ in range(60))
    loop("same pinned batch 60x", One())
t0 = time.time(); rep = T.train(eng, corpus, 0.025, 2); print("trainer.train", rep.steps, "steps", rep.seconds / rep.steps * 1e3, "ms/step", rep.device_ms)
