#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^W0" | tail -2
