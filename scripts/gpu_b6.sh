#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q 2>&1 | tail -3
for nn in 64 32; do
timeout 300 python bench.py --neg-sharing tile --tile-negatives $nn --no-baseline --no-fit --no-e2e --steps 20 --warmup 5 > gpurun_out/bench_tile_nn$nn.json 2> gpurun_out/bench_tile_nn$nn.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_tile_nn$nn.json").read())
print("nn=$nn", {k:d[k] for k in ("value","ms_per_step","loss_per_pair_first","loss_per_pair_last","max_abs_dot")})
PY
done
timeout 300 python bench.py --neg-sharing tile --dim 64 --tile-negatives 32 --no-baseline --no-fit --no-e2e --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('d=64 nn=32', d['value'], d['ms_per_step'])"
