for k in 2 0; do
  echo "MPLX_KERNEL=$k"
  MPLX_KERNEL=$k timeout 600 python tools/batch_queries.py --queries 4096 --max-expand 300 --ref-queries 0 2>/dev/null | tail -1 | python -c "
import json,sys
s=sys.stdin.readline().strip(); s=s[s.index('{'):]
d=json.loads(s); print(round(d['value']), d['seconds'], d['phase_seconds'], d['release_seconds_not_in_value'])"
done
timeout 300 python -m pytest tests/test_multi_query_gpu.py tests/test_planner_e2e_gpu.py -m gpu -x -q 2>&1 | tail -2
