#!/bin/bash
# Regenerates the round's tracked measurement artefacts on the GPU box (run through gpurun):
# outputs land in gpurun_out/ and are copied/summarised into profiles/ afterwards.
R=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${R}_pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/${R}_ref.err | tail -1 > gpurun_out/${R}_bench_reference_line.json
timeout 900 python bench.py 2>gpurun_out/${R}_bench.err | tail -1 > gpurun_out/${R}_bench_line.json
for w in cfg2 cfg3 cfg4; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 2>gpurun_out/${R}_bench_${w}.err | tail -1 > gpurun_out/${R}_bench_${w}.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-multi-query --no-replay > gpurun_out/${R}_ncu_list.log 2>&1
for f in gpurun_out/${R}_bench_line.json gpurun_out/${R}_bench_reference_line.json gpurun_out/${R}_bench_cfg*.json; do
  python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    keys=("value","ms_per_step","impl")
    print(sys.argv[1], {k:d.get(k) for k in keys}, "e2e", (d.get("e2e") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "mq", (d.get("multi_query") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
P
done
ls -la gpurun_out/ | tail -12
