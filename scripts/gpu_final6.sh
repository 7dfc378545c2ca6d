#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_quality.py -x -q -m gpu -k "golden or quality or async or zero_pair" 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^W0" | tail -2
timeout 400 python bench.py --no-baseline --no-tile > gpurun_out/bench_final6.json 2> gpurun_out/bench_final6.err; tail -2 gpurun_out/bench_final6.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_final6.json").read())
print(d["value"], d["e2e"]["value"], d["e2e_fit"])
PY
