import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from glint_word2vec_b200 import _C
from glint_word2vec_b200.data.corpus import chunk_encoded, iter_steps
from glint_word2vec_b200.data.sampler import build_alias, zipf_counts
from glint_word2vec_b200.models.engine import EngineOptions, ShardEngine
from glint_word2vec_b200.models.sgns import SGNSConfig
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
t0 = time.perf_counter(); pool = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(20)]; print("20 pin_memory allocs: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
batches = list(iter_steps(corpus, B))
ops = eng._cuda
def run(name, pos_fn, tok_fn):
    torch.cuda.synchronize(); acc = {"stage": 0.0, "launch": 0.0, "stats": 0.0, "wait": 0.0}; pend = []; t00 = time.perf_counter()
    for i, b in enumerate(batches):
        tk = tok_fn(i, b)
        a = time.perf_counter(); si, t = ops.stage_tokens(tk, b.sent_id); c = time.perf_counter()
        st = ops.train_step_staged(si, t, pos_fn(i, b), 0, 0.02); d = time.perf_counter()
        pend.append(st)
        if len(pend) > 3: pend.pop(0).cpu()
        e = time.perf_counter()
        acc["stage"] += c - a; acc["launch"] += d - c; acc["wait"] += e - d
    torch.cuda.synchronize(); tot = time.perf_counter() - t00
    print(name, "total %.3f ms/step" % (tot / len(batches) * 1e3), {k: round(v / len(batches) * 1e3, 3) for k, v in acc.items()}, flush=True)
for rep in range(2):
    run("real steps, real pos0      ", lambda i, b: b.raw_pos0, lambda i, b: b.tokens)
    run("real steps, pos0 = 0       ", lambda i, b: 0, lambda i, b: b.tokens)
    run("same tokens, real pos0     ", lambda i, b: b.raw_pos0, lambda i, b: batches[0].tokens if len(batches[0].tokens) == len(b.tokens) else b.tokens)
    run("same tokens, pos0 = 0      ", lambda i, b: 0, lambda i, b: batches[0].tokens if len(batches[0].tokens) == len(b.tokens) else b.tokens)
    run("real steps, pos0 = i*131072", lambda i, b: i * 131072, lambda i, b: b.tokens)
