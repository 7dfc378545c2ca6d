cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q -k "not probes" 2>&1 | tail -5
for nn in 32 64; do for d in 512 64; do
  timeout 300 python bench.py --steps 20 --warmup 3 --neg-sharing tile --tile-negatives $nn --no-e2e --dim $d 2>&1 | tail -1 | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('tile nn=$nn d=$d', round(r['value']/1e6,1),'Mpairs/s', round(r['ms_per_step'],4),'ms')"
done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgns_tile_kernel -s 3 -c 1 -o gpurun_out/prof_tile5_d512 python bench.py --steps 3 --warmup 2 --neg-sharing tile --no-e2e > gpurun_out/ncu_tile5_d512.log 2>&1
