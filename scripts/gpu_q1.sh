cd $GRAFT_REPO_ROOT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tile.py -x -q -k "damping" 2>&1 | tail -5
timeout 1500 python -m pytest tests/test_gpu_quality.py -q 2>&1 | tail -25
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgns_tile_kernel -s 3 -c 1 -o gpurun_out/prof_tile6_d64 python bench.py --steps 3 --warmup 2 --neg-sharing tile --tile-negatives 32 --no-e2e --dim 64 > gpurun_out/ncu_tile6_d64.log 2>&1
