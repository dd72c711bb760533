#!/bin/bash
# round 2, first device run: parity of the fixed-point kernel + named-shape parity, bench variants, ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests/test_expand_parity_gpu.py tests/test_full_shape_parity_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench_fx.json 2> gpurun_out/r02a_bench_fx.err; tail -c 3000 gpurun_out/r02a_bench_fx.json
MPLX_FX_UNR=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02a_bench_fx_unr4.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --kernel 2 2>/dev/null | tail -1 > gpurun_out/r02a_bench_reg.json
for w in cfg2 cfg3; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02a_bench_${w}_fx.json
done
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --kernel 4 2>/dev/null | tail -1 > gpurun_out/r02a_bench_cfg3_deal.json
for f in gpurun_out/r02a_bench_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1], "ms", d.get("ms_per_step"), "value", d.get("value"), "frac", (d.get("roofline") or {}).get("frac"), "kern", (d.get("roofline") or {}).get("kernel"), "parity", d.get("parity_checked"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
P
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_fx_kernel --launch-skip 2 --launch-count 1 -f \
  -o gpurun_out/prof_r02a_fx python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02a_ncu_fx.log 2>&1
ls -la gpurun_out/*.ncu-rep
