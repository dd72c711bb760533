set -x
timeout 900 python -m pytest tests/test_expand_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for k in 0 4; do for r in 0; do echo "kernel $k"; MPLX_DEAL_ROUNDS=$r timeout 300 python bench.py --kernel $k --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['ms_per_step'], d['value'])"; done; done
for r in 1 2 4 8; do echo "deal rounds $r"; MPLX_DEAL_ROUNDS=$r timeout 300 python bench.py --kernel 4 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['ms_per_step'], d['value'])"; done
for w in cfg3 cfg4; do for k in 0 4; do echo "$w kernel $k"; timeout 300 python bench.py --workload $w --kernel $k --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['ms_per_step'], d['value'])"; done; done
