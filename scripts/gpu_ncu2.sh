#!/bin/bash
# ncu captures of the shipped tile kernel (d = 512, NN = 64) and of the NN select kernel (one GPU; numbers under ncu are never bench values)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sgns_tile_kernel -s 3 -c 1 -f -o gpurun_out/prof_tile7_d512 python bench.py --steps 3 --warmup 2 --windows 2 --neg-sharing tile --no-e2e --no-fit --no-baseline > gpurun_out/ncu_tile7_d512.log 2>&1
tail -2 gpurun_out/ncu_tile7_d512.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nn_select_kernel -s 3 -c 1 -f -o gpurun_out/prof_nn_select python benchmarks/bench_nn.py --gpus 1 --iters 2 > gpurun_out/ncu_nn_select.log 2>&1
tail -2 gpurun_out/ncu_nn_select.log | cut -c1-200
ls -la gpurun_out/prof_tile7_d512.ncu-rep gpurun_out/prof_nn_select.ncu-rep
