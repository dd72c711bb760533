#!/bin/bash
# e2e A/B: packed tests, then bench e2e legs for a list of environment variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_expand_parity_gpu.py tests/test_multi_query_gpu.py -m gpu -x -q -k "packed or multi" 2>&1 | tail -4
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-multi-query --no-replay 2>gpurun_out/e2e_$i.err | tail -1 > gpurun_out/e2e_$i.json
  python - gpurun_out/e2e_$i.json "$envs" <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print(sys.argv[2], "| value", round(d["value"]/1e8,3), "e2e", round(d["e2e"]["value"]/1e8,3), "ms", round(d["e2e"]["ms_per_step"],3), "state", round(d["e2e_state_records"]["value"]/1e7,3), "full", round(d["e2e_full_contract"]["value"]/1e7,3))
P
done
