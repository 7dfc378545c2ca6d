#!/bin/bash
# quality grid + many-negatives tests + bench validation (1 GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "oracle or rejected or table_mode" 2>&1 | tail -5
timeout 1500 python scripts/gpu_q2.py 2>&1 | grep -v "^W0" > gpurun_out/q2_grid.jsonl
cat gpurun_out/q2_grid.jsonl | cut -c1-260
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
tail -3 gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json
