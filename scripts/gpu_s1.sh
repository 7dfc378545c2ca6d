#!/bin/bash
# compute-sanitizer over the round-2 kernels (1 GPU): tile kernel (single + loopback exchange), nn_select, pair kernel >7 negatives
cd "$(dirname "$0")/.."
TAG=tile FILES=tests/test_gpu_tile.py FILTER="one_tile_with_duplicates and 64-32 or loopback and 2-64" TOOLS="memcheck racecheck synccheck" LIMIT=500 bash scripts/sanitize.sh
TAG=nn FILES=tests/test_gpu_ops.py FILTER="nn_select_matches and 64-1-10" TOOLS="memcheck racecheck" LIMIT=400 bash scripts/sanitize.sh
TAG=pairs FILES=tests/test_gpu_ops.py FILTER="single_matches_oracle and 64-4-16" TOOLS="memcheck racecheck" LIMIT=400 bash scripts/sanitize.sh
timeout 600 python scripts/gpu_p1.py 2>&1 | grep -v "^W0" > gpurun_out/p1_fit_profile.txt; head -70 gpurun_out/p1_fit_profile.txt
