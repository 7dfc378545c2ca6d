#!/bin/bash
# 8-GPU job (charged 8x): keep it short.  bench N=8, a subset of the world-8 tests, NN bench, the other BASELINE configs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29801 bench.py --gpus 8 --steps 20 --warmup 5 --windows 6 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err
tail -2 gpurun_out/bench_r2_n8.err | cut -c1-300; cat gpurun_out/bench_r2_n8.json
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "test_fused_multi_matches_oracle and 8-512-1-1-16 or many_negatives and 8-300 or tile_kernel_column_shards and 8-512-64 or nn_row_shard and 8-300" 2>&1 | tail -6
timeout 300 python benchmarks/bench_nn.py --gpus 8 > gpurun_out/nn_8gpu.json 2> gpurun_out/nn_8gpu.err; tail -2 gpurun_out/nn_8gpu.err | cut -c1-300; cat gpurun_out/nn_8gpu.json
timeout 300 $TR --master-port 29802 bench.py --gpus 8 --vocab 1000000 --dim 300 --steps 20 --warmup 5 --windows 5 --no-fit > gpurun_out/bench_r2_n8_1m300.json 2> gpurun_out/bench_r2_n8_1m300.err
tail -2 gpurun_out/bench_r2_n8_1m300.err | cut -c1-300; cat gpurun_out/bench_r2_n8_1m300.json
timeout 420 $TR --master-port 29803 bench.py --gpus 8 --vocab 80000000 --dim 300 --neg 10 --steps 20 --warmup 5 --windows 4 --no-fit --no-baseline --no-selfcheck > gpurun_out/bench_r2_n8_80m300.json 2> gpurun_out/bench_r2_n8_80m300.err
tail -2 gpurun_out/bench_r2_n8_80m300.err | cut -c1-300; cat gpurun_out/bench_r2_n8_80m300.json
