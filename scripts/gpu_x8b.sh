#!/bin/bash
# second 8-GPU job: tile kernel with the coalesced exchange payload + the final bench line at N = 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29811 bench.py --gpus 8 --steps 20 --warmup 5 --windows 6 > gpurun_out/bench_r2_n8b.json 2> gpurun_out/bench_r2_n8b.err
tail -2 gpurun_out/bench_r2_n8b.err | cut -c1-300; cat gpurun_out/bench_r2_n8b.json
timeout 150 python -m pytest tests/test_gpu_multi.py -x -q -k "tile_kernel_column_shards and 8-512-64 or exchange_formats and 8-512" 2>&1 | tail -4
