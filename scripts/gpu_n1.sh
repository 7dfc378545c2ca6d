#!/bin/bash
# tile kernel regression + NN select tests + benches (1 GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile.py -x -q 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "nn_select or topk or inference" 2>&1 | tail -15
timeout 600 python bench.py --neg-sharing tile --no-baseline --no-fit --no-e2e --steps 20 --warmup 5 > gpurun_out/bench_tileonly.json 2> gpurun_out/bench_tileonly.err
tail -2 gpurun_out/bench_tileonly.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_tileonly.json").read())
print({k:d[k] for k in ("value","ms_per_step","loss_per_pair_first","loss_per_pair_last","max_abs_dot")})
PY
timeout 600 python benchmarks/bench_nn.py --gpus 1 > gpurun_out/nn_1gpu.json 2> gpurun_out/nn_1gpu.err; tail -2 gpurun_out/nn_1gpu.err; cat gpurun_out/nn_1gpu.json
timeout 600 python benchmarks/bench_nn.py --gpus 1 --dim 300 --queries 64 > gpurun_out/nn_1gpu_d300.json 2>> gpurun_out/nn_1gpu.err; cat gpurun_out/nn_1gpu_d300.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
tail -2 gpurun_out/bench_r2b.err; cat gpurun_out/bench_r2b.json
