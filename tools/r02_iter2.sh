#!/bin/bash
# tuning iteration: parity of the expansion kernels, bench of the four workloads, one ncu capture of the headline kernel.
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_expand_parity_gpu.py tests/test_full_shape_parity_gpu.py -m gpu -x -q 2>&1 | tail -5
run() { # name, env..., -- bench args
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-multi-query --no-replay $BARGS 2>gpurun_out/${TAG}_${name}.err | tail -1 > gpurun_out/${TAG}_${name}.json
  python - gpurun_out/${TAG}_${name}.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1], "ms", round(d.get("ms_per_step"),4), "frac", round(d["roofline"]["frac"],4), d["roofline"]["kernel"], d["config"]["workload"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
P
}
BARGS="" run head X=1
BARGS="" run head2 X=1
BARGS="--workload cfg2" run cfg2 X=1
BARGS="--workload cfg3" run cfg3 X=1
BARGS="--workload cfg4" run cfg4 X=1
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_fx --launch-skip 2 --launch-count 1 -f \
  -o gpurun_out/prof_${TAG}_fx python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-multi-query --no-replay > gpurun_out/${TAG}_ncu_fx.log 2>&1
ncu -i gpurun_out/prof_${TAG}_fx.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_fx_raw.csv 2>/dev/null; ncu -i gpurun_out/prof_${TAG}_fx.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_${TAG}_fx_sass.csv 2>/dev/null; rm -f gpurun_out/prof_${TAG}_fx.ncu-rep
fi
