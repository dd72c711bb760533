#!/bin/bash
# full device check: every -m gpu test, smoke, the default bench (both arms)
TAG=${1:-full}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 1500 gpurun_out/${TAG}_bench_ref.json
