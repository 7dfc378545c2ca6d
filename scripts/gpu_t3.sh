#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python scripts/gpu_t3.py 10000000 512 '[[64,32,null,1.0],[64,32,1.0,0.4],[64,32,1.0,0.2],[64,32,null,0.4],[64,8,null,1.0],[32,32,1.0,0.4]]' 2>&1 | grep -v "^W0" > gpurun_out/t3_dyn.jsonl; cut -c1-600 gpurun_out/t3_dyn.jsonl
bash scripts/gpu_s1.sh
