#!/bin/bash
# fx kernel tuning sweep: parity first, then MPLX_FX_UNR x MPLX_FX_MINB on the headline workload.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_expand_parity_gpu.py tests/test_full_shape_parity_gpu.py -m gpu -x -q 2>&1 | tail -5
for unr in 4 8; do for minb in 4 5 6; do
  MPLX_FX_UNR=$unr MPLX_FX_MINB=$minb timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/sweep_${unr}_${minb}.json
  python - gpurun_out/sweep_${unr}_${minb}.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1], "ms", round(d.get("ms_per_step"),4), "frac", round(d["roofline"]["frac"],4), d["roofline"]["kernel"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
P
done; done
