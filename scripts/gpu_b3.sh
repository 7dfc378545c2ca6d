cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_tile.py -x -q -k "not probes" 2>&1 | tail -15
