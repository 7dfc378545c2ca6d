#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -q -m gpu 2>&1 | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^W0" | tail -3
