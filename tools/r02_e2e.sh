#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for mode in direct staged; do
  if [ $mode = staged ]; then export MPLX_PACK_STAGED=1; else unset MPLX_PACK_STAGED; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-multi-query --no-replay 2>gpurun_out/e2e_$mode.err | tail -1 > gpurun_out/e2e_$mode.json
  python - gpurun_out/e2e_$mode.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print(sys.argv[1], "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["e2e"]["ms_per_step"], "GB/s", d["e2e"]["d2h_gbs"], "state", d["e2e_state_records"]["value"])
P
done
unset MPLX_PACK_STAGED
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-multi-query 2>gpurun_out/replay.err | tail -1 > gpurun_out/replay.json
python - <<'P'
import json
d=json.loads(open("gpurun_out/replay.json").read().strip().split("\n")[-1])
print(json.dumps(d.get("replay"), indent=0)[:1500])
P
