#!/bin/bash
py() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$1', d['ms_per_step'], (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('ms_per_step'), (d.get('e2e_state_records') or {}).get('ms_per_step'))"; }
for lg in 20 21 22 23; do MPLX_PACK_CHUNK_LOG2=$lg python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-multi-query --no-replay 2>/dev/null | tail -1 | py lg$lg; done
