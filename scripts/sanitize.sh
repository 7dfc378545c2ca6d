#!/usr/bin/env bash
# Race / memory / sync checking of the single-GPU kernels.  Run on a GPU box:
#   gpurun --timeout 1200 -- scripts/sanitize.sh
# Intended races (documented): fp32 RED.128 / TMA-reduce adds on shared embedding rows (Hogwild) and
# stale row reads between concurrently processed pairs.  Forbidden: out-of-bounds accesses, shared-memory
# hazards inside a warp's stage ring, barrier misuse -- memcheck / racecheck / synccheck must stay clean.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# FILES / TOOLS / FILTER / LIMIT / TAG can be narrowed from the environment when GPU time is short
T="${FILES:-tests/test_gpu_ops.py} -m gpu -q -x -k '${FILTER:-zero_pair or subsample or zipf or init_syn0 or inference}'"
for tool in ${TOOLS:-memcheck synccheck racecheck}; do
  echo "== compute-sanitizer --tool $tool"
  eval timeout ${LIMIT:-900} compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $T > gpurun_out/sanitize_${TAG:-r2}_$tool.log 2>&1
  echo "exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitize_${TAG:-r2}_$tool.log | tail -2 | tr '\n' ' ')"
done
