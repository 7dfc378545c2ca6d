#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "tile_kernel_column_shards and 2- or exchange_formats and 2-" 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 20 --warmup 5 --windows 6 > gpurun_out/bench_final_n2.json 2> gpurun_out/bench_final_n2.err
tail -2 gpurun_out/bench_final_n2.err | cut -c1-200; cat gpurun_out/bench_final_n2.json
