#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python scripts/gpu_t1.py 1000000 512 300 '[["tile",64,64],["tile",128,64],["tile",1000,64],["tile",64,32],["tile",128,32],["tile",32,64]]' 2>&1 | grep -v "^W0" > gpurun_out/t2_dyn.jsonl; cut -c1-900 gpurun_out/t2_dyn.jsonl
python - <<'PY' 2>&1 | grep -v "^W0"
import sys, json, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_gpu_quality as q
for nn in (32, 64):
    for cap in (32.0, 128.0):
        for st in (8192, 131072):
            rep, rec, vec = q._fit({"neg_sharing": "tile", "tile_negatives": nn, "step_tokens": st, "hot_row_cap": cap})
            print(json.dumps({"nn": nn, "cap": cap, "step": st, "loss": round(rep["loss_per_pair"], 4), "recall": rec,
                              "maxnorm": round(float(np.linalg.norm(vec, axis=1).max()), 2)}), flush=True)
PY
