#!/bin/bash
# A/B of environment variants on one workload: tools/r02_ab.sh TAG "bench args" "VAR=val ..." "VAR=val ..." ...
TAG=$1; BARGS=$2; shift 2
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-multi-query --no-replay $BARGS 2>gpurun_out/${TAG}_$i.err | tail -1 > gpurun_out/${TAG}_$i.json
  python - gpurun_out/${TAG}_$i.json "$envs" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[2], "| ms", round(d.get("ms_per_step"),4), "frac", round(d["roofline"]["frac"],4), d["config"]["workload"], "parity", d.get("parity_checked"))
except Exception as e:
    print(sys.argv[2], "unreadable:", e)
P
done
