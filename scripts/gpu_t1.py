"""Tile-vs-pair learning dynamics on the benchmark-like stream (i.i.d. Zipf tokens): loss / max|dot| trajectories."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from glint_word2vec_b200.data.sampler import build_alias, zipf_counts, zipf_tokens
from glint_word2vec_b200.models.engine import EngineOptions, ShardEngine
from glint_word2vec_b200.models.sgns import SGNSConfig

V, D, T, STEPS = int(sys.argv[1]), int(sys.argv[2]), 131072, int(sys.argv[3])
counts = zipf_counts(V, 10 ** 9, 1.0)
alias = build_alias(counts.astype(np.float64))
dev = torch.device("cuda", 0)
toks = [zipf_tokens(alias, T, seed=100 + s) for s in range(8)]
sid = (np.arange(T) // 1000).astype(np.int32)
CONFIGS = json.loads(sys.argv[4]) if len(sys.argv) > 4 else [["pair", 32.0, 32], ["pair", 0.0, 32], ["tile", 32.0, 32], ["tile", 0.0, 32], ["tile", 32.0, 64]]
for mode, cap, nn in CONFIGS:
    cfg = SGNSConfig(V, D, 5, 5, seed=1, neg_sharing=mode, tile_negatives=nn)
    eng = ShardEngine(cfg, device=dev, options=EngineOptions(subsample_mode="word2vec", subsample_ratio=1e-4, hot_row_cap=cap))
    eng.init_weights(); eng.set_noise(counts)
    traj = []
    for s in range(STEPS):
        st = eng.train_step(toks[s % 8], sid, s * T, 0, 0.025)
        if s % 25 == 24 or s == 0:
            st = st.cpu()
            traj.append((s, round(float(st[1] / max(st[0], 1)), 4), round(float(st[2]), 2)))
    print(json.dumps({"mode": mode, "cap": cap, "nn": nn, "V": V, "d": D, "traj": traj,
                      "max_norm0": float(eng.syn0.norm(dim=1).max()), "max_norm1": float(eng.syn1.norm(dim=1).max())}), flush=True)
    eng.destroy(); del eng; torch.cuda.empty_cache()
