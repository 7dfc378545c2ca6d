"""Diagnostic: tile kernel vs oracle over a grid of (loopback world, d, NN, tokens)."""
import os, sys, subprocess, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def one(world, d, nn, t, grid=0, steps=1):
    import torch
    if world > 1:
        os.environ["GW2V_LOOPBACK_WORLD"] = str(world)
    if grid:
        os.environ["GW2V_TILE_GRID"] = str(grid)
    from glint_word2vec_b200.data.sampler import zipf_counts
    from glint_word2vec_b200.models import sgns
    from glint_word2vec_b200.models.engine import EngineOptions, ShardEngine
    from glint_word2vec_b200.models.sgns import SGNSConfig
    dev = torch.device("cuda", 0)
    v = 200000
    cfg = SGNSConfig(v, d, 5, 5, seed=7, neg_sharing="tile", tile_negatives=nn)
    eng = ShardEngine(cfg, device=dev, options=EngineOptions(subsample_mode="reference", hot_row_cap=0))
    eng.init_weights(); eng.set_noise(zipf_counts(v, 10 ** 7, 0.6))
    g = torch.Generator().manual_seed(0)
    syn1 = torch.randn(v, eng.shard.cols, generator=g) * (0.5 / d ** 0.5)
    syn0 = torch.randn(v, eng.shard.cols, generator=g) * (0.5 / d ** 0.5)
    eng.syn0, eng.syn1 = syn0.to(dev), syn1.to(dev)
    rng = np.random.default_rng(4)
    ref0, ref1 = syn0.clone(), syn1.clone()
    pos = 0
    for _ in range(steps):
        tokens = rng.choice(v, size=t, replace=False).astype(np.int32)
        sid = (np.arange(t) // 29).astype(np.int32)
        st = sgns.sgns_minibatch_reference(ref0, ref1, cfg, eng.alias, tokens, sid, pos, 0, 0.002)
        stats = eng.train_step(tokens, sid, pos, 0, 0.002).cpu()
        pos += t
    got0, got1 = eng.syn0.cpu(), eng.syn1.cpu()
    r0, r1 = ref0 - syn0, ref1 - syn1
    e0 = float((got0 - syn0 - r0).norm() / r0.norm()); e1 = float((got1 - syn1 - r1).norm() / r1.norm())
    # split syn1 error: negative rows vs context rows
    tn = sgns.tile_negatives(cfg, eng.alias, pos - t, np.arange((t + 127) // 128), 0).reshape(-1).astype(np.int64)
    isneg = torch.zeros(v, dtype=torch.bool); isneg[torch.from_numpy(np.unique(tn))] = True
    d1 = got1 - syn1 - r1
    en = float(d1[isneg].norm() / r1[isneg].norm()); ec = float(d1[~isneg].norm() / max(float(r1[~isneg].norm()), 1e-30))
    print("RESULT", json.dumps({"world": world, "d": d, "nn": nn, "t": t, "grid": grid, "steps": steps, "e0": round(e0, 4), "e1": round(e1, 4), "e1_neg": round(en, 4),
                                "e1_ctx": round(ec, 4), "pairs_ok": int(stats[0]) == st.pairs}), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(*[int(a) for a in sys.argv[1:7]])
    else:
        cases = [(1, 512, 32, 40000, 0, 2), (1, 512, 32, 3000, 2, 1), (1, 256, 32, 20000, 8, 2), (1, 128, 32, 3000, 4, 1), (1, 128, 64, 3000, 4, 1), (1, 128, 64, 3000, 12, 3), (8, 128, 64, 3000, 4, 3), (1, 512, 64, 3000, 2, 1), (1, 512, 64, 40000, 0, 2), (4, 100, 32, 40000, 0, 2), (1, 300, 64, 9000, 3, 2)]
        for c in cases:
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + [str(x) for x in c], capture_output=True, text=True, timeout=300)
            out = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            print(out[0] if out else ("CRASH %s " % (c,) + r.stderr[-300:]), flush=True)
