#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -q -m gpu 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^W0" | tail -3
timeout 600 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -2 gpurun_out/bench_final_n1.err | cut -c1-200; cat gpurun_out/bench_final_n1.json
