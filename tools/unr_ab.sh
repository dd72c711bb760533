#!/bin/bash
run() { python bench.py "$@" --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print(' ', d['roofline']['kernel'], d['ms_per_step'], round(d['value']/1e6,1), 'M/s')"; }
echo "UNR4"; run --steps 20 --warmup 3
echo "UNR6"; MPLX_UNR6=1 run --steps 20 --warmup 3
echo "UNR8"; MPLX_UNR8=1 run --steps 20 --warmup 3
echo "cfg2 UNR4"; run --workload cfg2 --steps 10 --warmup 3
echo "cfg2 UNR8"; MPLX_UNR8=1 run --workload cfg2 --steps 10 --warmup 3
MPLX_UNR8=1 timeout 300 python -m pytest tests/test_expand_parity_gpu.py -m gpu -x -q -k "headline or cfg2" 2>&1 | tail -2
