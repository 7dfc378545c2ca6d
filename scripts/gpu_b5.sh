#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q -k "loopback" 2>&1 | tail -3
timeout 900 python scripts/demo_stream.py --tokens 200000000 --vocab 1000000 --dim 128 2> gpurun_out/demo_stream.err | grep "^{" > gpurun_out/demo_stream.json; tail -2 gpurun_out/demo_stream.err | cut -c1-200; cat gpurun_out/demo_stream.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-baseline > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -2 gpurun_out/bench_r2d.err | cut -c1-200; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2d.json").read())
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["e2e_fit"], d["tile"]["value"])
PY
