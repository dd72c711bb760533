#!/usr/bin/env python
"""BASELINE.json config 5 from the command line (the same run `bench.py --workload cfg5` times):

  python tools/batch_queries.py [--queries 4096] [--max-expand 1000] [--eps 2] [--cells 512]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/batch_queries.py ...

Prints one JSON line on rank 0 (see cfg5_bench.run)."""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--max-expand", type=int, default=1000)
    ap.add_argument("--eps", type=float, default=2.0)
    ap.add_argument("--cells", type=int, default=512)
    ap.add_argument("--min-dist", type=float, default=20.0)
    ap.add_argument("--ref-queries", type=int, default=256)
    ap.add_argument("--repeat", type=int, default=2)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    import torch
    import torch.distributed as dist

    import cfg5_bench
    import scenarios as S
    from motion_primitive_library_b200 import sharding

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = S.cfg3() if args.cells == 512 else S.scaled(S.cfg3(), args.cells)
    # rank 0 builds the map, every rank receives its replica (the only set-up collective: SURVEY.md §8e)
    grid = sharding.broadcast_array(sc.grid() if rank == 0 else None, src=0)
    out = cfg5_bench.run(sc, grid, local, n_queries=args.queries, max_expand=args.max_expand, min_dist=args.min_dist, eps=args.eps,
                         ref_queries=args.ref_queries, repeat=args.repeat)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
