"""Quality grid on one B200: planted corpus, CPU reference mini-batches vs the GPU kernels (numbers for tests/test_gpu_quality.py)."""
import json, sys, time
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_gpu_quality as q

def run(tag, cfg):
    t = time.time()
    rep, rec, vec = q._fit(cfg)
    out = {"tag": tag, "cfg": cfg, "loss": rep["loss_per_pair"], "recall": rec,
           "maxnorm": float(np.linalg.norm(vec, axis=1).max()) if np.isfinite(vec).all() else float("inf"),
           "finite": bool(np.isfinite(vec).all()), "sec": round(time.time() - t, 1)}
    print(json.dumps(out), flush=True)

for sub in ("word2vec", "reference"):
    run("cpu_ref", {"device": "cpu", "subsample_mode": sub})
    for mode in ("pair", "tile"):
        for st in (8192, 131072):
            for cap in (32.0, 0.0):
                run("gpu", {"neg_sharing": mode, "step_tokens": st, "hot_row_cap": cap, "subsample_mode": sub})
