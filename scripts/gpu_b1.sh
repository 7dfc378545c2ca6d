set -x
cd $GRAFT_REPO_ROOT
for ns in pair tile; do
  timeout 300 python bench.py --steps 20 --warmup 3 --neg-sharing $ns --no-e2e 2>&1 | tail -1 > gpurun_out/b1_${ns}_d512.json
  timeout 300 python bench.py --steps 20 --warmup 3 --neg-sharing $ns --no-e2e --dim 64 2>&1 | tail -1 > gpurun_out/b1_${ns}_d64.json
done
timeout 300 python bench.py --steps 20 --warmup 3 --neg-sharing tile --tile-negatives 32 --no-e2e 2>&1 | tail -1 > gpurun_out/b1_tile32_d512.json
timeout 300 python bench.py --steps 20 --warmup 3 --neg-sharing tile --tile-negatives 32 --no-e2e --dim 64 2>&1 | tail -1 > gpurun_out/b1_tile32_d64.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgns_tile_kernel -s 3 -c 1 -o gpurun_out/prof_tile_d512 python bench.py --steps 3 --warmup 2 --neg-sharing tile --no-e2e > gpurun_out/ncu_tile_d512.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgns_tile_kernel -s 3 -c 1 -o gpurun_out/prof_tile_d64 python bench.py --steps 3 --warmup 2 --neg-sharing tile --no-e2e --dim 64 > gpurun_out/ncu_tile_d64.log 2>&1
cat gpurun_out/b1_*.json
