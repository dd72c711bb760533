#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_expand_parity_gpu.py tests/test_full_shape_parity_gpu.py tests/test_integration_gpu.py tests/test_multi_query_gpu.py -m gpu -x -q 2>&1 | tail -4
for k in 0 2 4; do
timeout 900 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --kernel $k 2>gpurun_out/cfg4_$k.err | tail -1 > gpurun_out/cfg4_$k.json
python - gpurun_out/cfg4_$k.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1], "ms", round(d.get("ms_per_step"),4), "frac", round(d["roofline"]["frac"],4), d["roofline"]["kernel"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
P
done
