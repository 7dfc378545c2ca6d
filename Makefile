# Build / test / bench entry points (the reference drives everything through sbt: `sbt assembly`,
# `sbt it:test`, build.sbt:42-82; here plain make + python).
PY ?= python

build:            ## compile _C.so (sm_100a kernels) and _host.so in-tree; no GPU needed
	$(PY) -m glint_word2vec_b200.build_ext

test: build       ## CPU tiers (includes the Gloo world_size=2 and the 15 golden scenarios)
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu: build   ## on a B200 box
	$(PY) -m pytest tests -q -m gpu

bench: build      ## headline metric, one GPU; use `torchrun` via bench.py --gpus N for more
	$(PY) bench.py --gpus 1

server:           ## stand-alone shard-server group (cf. `spark-submit --class glint.Main`)
	$(PY) -m glint_word2vec_b200.parallel.server -c configs/separate-server.json

sass:             ## Blackwell SASS evidence table
	scripts/sass_evidence.sh > profiles/sass_summary.md

.PHONY: build test test-gpu bench server sass
