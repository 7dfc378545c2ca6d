cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "tile_kernel_column_shards and (2-128-32 or 2-100-64)" 2>&1 | tail -8
for ns in pair tile; do
  timeout 300 python bench.py --gpus 2 --steps 20 --warmup 3 --neg-sharing $ns --tile-negatives 32 --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('$ns x2', round(r['value']/1e6,1),'Mpairs/s', round(r['ms_per_step'],4),'ms', 'wait_us', r.get('exposed_allreduce_us_per_step'))"
done
