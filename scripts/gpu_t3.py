"""Pair-pretrained weights -> tile mode: which damping keeps it stable?  (the bench.py scenario that exploded)"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from glint_word2vec_b200.data.sampler import build_alias, zipf_counts, zipf_tokens
from glint_word2vec_b200.models.engine import EngineOptions, ShardEngine
from glint_word2vec_b200.models.sgns import SGNSConfig
V, D, T = int(sys.argv[1]), int(sys.argv[2]), 131072
counts = zipf_counts(V, 200 * T, 1.0)
alias = build_alias(counts.astype(np.float64))
dev = torch.device("cuda", 0)
toks = [zipf_tokens(alias, T, seed=100 + s) for s in range(16)]
sid = (np.arange(T) // 1000).astype(np.int32)
def mk(mode, nn=64, cap=32.0):
    cfg = SGNSConfig(V, D, 5, 5, seed=1, neg_sharing=mode, tile_negatives=nn)
    e = ShardEngine(cfg, device=dev, options=EngineOptions(subsample_mode="word2vec", subsample_ratio=1e-4, hot_row_cap=cap))
    return e
pe = mk("pair"); pe.init_weights(); pe.set_noise(counts)
for s in range(200):
    st = pe.train_step(toks[s % 16], sid, s * T, 0, 0.025)
st = st.cpu(); print("pair pretrain: loss", float(st[1] / st[0]), "maxdot", float(st[2]), "norms", float(pe.syn0.norm(dim=1).max()), float(pe.syn1.norm(dim=1).max()), flush=True)
base0, base1 = pe.syn0.clone(), pe.syn1.clone()
for nn, cap, ns, nw in json.loads(sys.argv[3]):
    if ns is not None: os.environ["GW2V_TILE_NEG_SCALE"] = str(ns)
    else: os.environ.pop("GW2V_TILE_NEG_SCALE", None)
    os.environ["GW2V_TILE_NEG_WEIGHT"] = str(nw)
    te = mk("tile", nn, cap); te.syn0, te.syn1 = base0.clone(), base1.clone(); te.set_noise(counts)
    traj = []
    for s in range(120):
        st = te.train_step(toks[(200 + s) % 16], sid, (200 + s) * T, 0, 0.025)
        if s in (0, 4, 9, 19, 39, 79, 119):
            st = st.cpu(); traj.append((s, round(float(st[1] / max(st[0], 1)), 3), round(float(st[2]), 1)))
    print(json.dumps({"nn": nn, "cap": cap, "neg_scale": ns, "neg_weight": nw, "traj": traj,
                      "norms": [round(float(te.syn0.norm(dim=1).max()), 1), round(float(te.syn1.norm(dim=1).max()), 1)]}), flush=True)
    te.destroy(); del te; torch.cuda.empty_cache()
