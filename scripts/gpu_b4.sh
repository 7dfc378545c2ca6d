#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err
tail -2 gpurun_out/bench_r2c.err | cut -c1-300; cat gpurun_out/bench_r2c.json
timeout 600 python scripts/gpu_p1.py 2>&1 | grep -v "^W0" > gpurun_out/p2_fit_profile.txt; head -45 gpurun_out/p2_fit_profile.txt | cut -c1-180
