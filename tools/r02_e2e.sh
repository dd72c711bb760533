#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_expand_parity_gpu.py tests/test_multi_query_gpu.py -m gpu -x -q -k "packed or multi" 2>&1 | tail -4
for lg in 20 19 21; do
  MPLX_PACK_CHUNK_LOG2=$lg timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-multi-query --no-replay 2>gpurun_out/e2e_$lg.err | tail -1 > gpurun_out/e2e_$lg.json
  python - gpurun_out/e2e_$lg.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print(sys.argv[1], "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["e2e"]["ms_per_step"], "GB/s", d["e2e"]["d2h_gbs"], "state", d["e2e_state_records"]["value"], "full", d["e2e_full_contract"]["value"])
P
done
for k in 2 4; do MPLX_DEAL_UNR=$k timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('cfg4 unr', $k, d['ms_per_step'])"; done
