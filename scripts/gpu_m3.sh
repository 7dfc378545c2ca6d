#!/bin/bash
# 2-GPU job: multi tests (world 2), bench N=2 with selfcheck, NN bench N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -k "2-" 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 --windows 6 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err
tail -3 gpurun_out/bench_r2_n2.err; cat gpurun_out/bench_r2_n2.json
timeout 600 python benchmarks/bench_nn.py --gpus 2 > gpurun_out/nn_2gpu.json 2> gpurun_out/nn_2gpu.err; tail -3 gpurun_out/nn_2gpu.err; cat gpurun_out/nn_2gpu.json
