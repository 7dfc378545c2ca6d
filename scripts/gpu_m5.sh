#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 4 --steps 20 --warmup 5 --windows 6 > gpurun_out/bench_final_n4.json 2> gpurun_out/bench_final_n4.err
tail -2 gpurun_out/bench_final_n4.err | cut -c1-200; cat gpurun_out/bench_final_n4.json
timeout 120 python benchmarks/bench_nn.py --gpus 4 2>/dev/null | grep "^{" > gpurun_out/nn_4gpu.json; cat gpurun_out/nn_4gpu.json
