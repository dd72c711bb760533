#!/bin/bash
# Regenerates the round's tracked measurement artefacts on the GPU box (run through gpurun):
# outputs land in gpurun_out/ and are copied/summarised into profiles/ afterwards.
R=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/${R}_ref.err | tail -1 > gpurun_out/${R}_bench_reference_line.json
timeout 900 python bench.py 2>gpurun_out/${R}_bench.err | tail -1 > gpurun_out/${R}_bench_line.json
for w in cfg2 cfg3 cfg4; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${R}_bench_${w}.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_list.log 2>&1
timeout 900 python tools/batch_queries.py --queries 4096 --max-expand 300 --ref-queries 128 2>/dev/null | tail -1 > gpurun_out/${R}_cfg5.json
for f in gpurun_out/${R}_bench_line.json gpurun_out/${R}_bench_reference_line.json gpurun_out/${R}_bench_cfg*.json gpurun_out/${R}_cfg5.json; do
  python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    keys=("value","ms_per_step","impl")
    print(sys.argv[1], {k:d.get(k) for k in keys}, "e2e", (d.get("e2e") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "ref", (d.get("reference") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
P
done
ls -la gpurun_out/ | tail -15
