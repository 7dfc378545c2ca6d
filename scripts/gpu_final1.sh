#!/bin/bash
# final single-GPU pass: what the driver runs at round end (pytest -m gpu, smoke, bench) + the streaming demonstration
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^W0" | tail -3
timeout 600 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -2 gpurun_out/bench_final_n1.err | cut -c1-200; cat gpurun_out/bench_final_n1.json
timeout 600 python scripts/demo_stream.py --tokens 200000000 --vocab 1000000 --dim 128 2> gpurun_out/demo_stream.err | grep "^{" > gpurun_out/demo_stream.json; cat gpurun_out/demo_stream.json
