#!/bin/bash
# quick kernel check on the GPU box: parity of all kernels, then ms/step of the default and selected kernels
timeout 900 python -m pytest tests/test_expand_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
run() { python bench.py "$@" --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print(' ', d['roofline']['kernel'], d['ms_per_step'], round(d['value']/1e6,1), 'M/s')"; }
echo "512c_acc27 auto"; run --steps 20 --warmup 3
echo "512c_acc27 kernel 4"; run --steps 20 --warmup 3 --kernel 4
for w in cfg2 cfg3 cfg4; do echo "$w auto"; run --workload $w --steps 10 --warmup 3; done
echo "cfg3 kernel 2"; run --workload cfg3 --steps 10 --warmup 3 --kernel 2
