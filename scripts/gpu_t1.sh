#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python benchmarks/diag_tile_loopback.py 2>&1 | grep -v "^W0" | head -12
timeout 900 python scripts/gpu_t1.py 1000000 512 150 2>&1 | grep -v "^W0" > gpurun_out/t1_dyn.jsonl; cut -c1-700 gpurun_out/t1_dyn.jsonl
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "nn_select" 2>&1 | tail -4
