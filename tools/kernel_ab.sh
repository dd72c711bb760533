#!/bin/bash
# quick kernel check on the GPU box: parity of all kernels, then ms/step of selected kernels
timeout 200 python -m pytest tests/test_expand_parity_gpu.py -m gpu -x -q 2>&1 | tail -1
run() { python bench.py "$@" --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print(' ', d['roofline']['kernel'], d['ms_per_step'], round(d['value']/1e6,1), 'M/s')"; }
echo "512c_acc27 kernel 4"; run --steps 10 --warmup 3 --kernel 4
echo "cfg3 auto"; run --workload cfg3 --steps 5 --warmup 3
