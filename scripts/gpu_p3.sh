#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python scripts/gpu_p3.py 2>&1 | grep -v "^W0\|UserWarning\|copy_(" | tail -20
