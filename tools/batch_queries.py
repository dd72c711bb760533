#!/usr/bin/env python
"""BASELINE.json config 5: many (start, goal) queries on one 512^3 voxel map, JRK control, planned in
lock-step by MPL::MultiQueryPlanner (one device launch per iteration expands the current node of
every live query of the rank; host bookkeeping spread over the host cores) and sharded over ranks.

  python tools/batch_queries.py [--queries 4096] [--max-expand 300] [--cells 512]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/batch_queries.py ...

Prints one JSON line on rank 0: total node expansions / max-over-ranks wall seconds of the search
(map upload excluded), plus the same queries planned one by one by the reference planner
(oracle/_ref, one query per host thread) on a bounded sample, when that library is present."""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--max-expand", type=int, default=300)
    ap.add_argument("--cells", type=int, default=512)
    ap.add_argument("--min-dist", type=float, default=20.0)
    ap.add_argument("--ref-queries", type=int, default=256)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    import torch
    import torch.distributed as dist

    from motion_primitive_library_b200 import planner, sharding
    from motion_primitive_library_b200 import scenarios as S

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = S.cfg3() if args.cells == 512 else S.scaled(S.cfg3(), args.cells)
    # rank 0 builds the map, every rank receives its replica (the only set-up collective: SURVEY.md §8e)
    grid = sharding.broadcast_array(sc.grid() if rank == 0 else None, src=0)
    # start/goal pairs: free cell centres at rest, at least min_dist apart
    pts = sc.frontier(4 * args.queries, seed=11, max_steps=0)["pos"]
    rng = np.random.default_rng(5)
    pairs = []
    while len(pairs) < args.queries:
        i, j = rng.integers(0, len(pts), 2)
        if np.abs(pts[i] - pts[j]).max() >= min(args.min_dist, 0.4 * args.cells * sc.res):
            pairs.append((i, j))
    q = np.zeros(args.queries, dtype=[("start", planner.WAYPOINT_DTYPE), ("goal", planner.WAYPOINT_DTYPE)])
    for k, (i, j) in enumerate(pairs):
        q["start"]["pos"][k], q["goal"]["pos"][k] = pts[i], pts[j]

    def make(start, goal):
        a = planner.make_args(3, sc.control, grid, sc.dim_cells, sc.origin, sc.res, sc.U, start=dict(pos=start), goal=dict(pos=goal),
                              v_max=sc.v_max, a_max=sc.a_max, T=sc.T, w=sc.w, max_num=args.max_expand)
        a.device = local
        return a

    def run_slice(mine):
        if len(mine) == 0:
            return np.zeros(0, dtype=[("valid", "i4"), ("cost", "f8"), ("expanded", "i4"), ("n_closed", "i4"), ("n_actions", "i4")]), \
                dict(expansions=0, iterations=0, seconds_max=0.0, t_pop_max=0.0, t_device_max=0.0, t_relax_max=0.0, t_release_max=0.0)
        res, tot = planner.plan_batch(make(mine["start"]["pos"][0], mine["goal"]["pos"][0]), mine["start"], mine["goal"])
        return res, dict(expansions=tot["nodes"], iterations=tot["iterations"], seconds_max=tot["seconds"], t_pop_max=tot["t_pop"],
                         t_device_max=tot["t_device"], t_relax_max=tot["t_relax"], t_release_max=tot["t_release"])

    if world > 1:
        dist.barrier()
    res, cnt = sharding.run_sharded(q, run_slice)
    if rank == 0:
        line = {"workload": f"cfg5: {args.queries} start/goal pairs, {sc.name}, max {args.max_expand} expansions/query",
                "n_gpus": world, "value": cnt["expansions"] / cnt["seconds_max"], "unit": "expansions/s",
                "expansions": int(cnt["expansions"]), "seconds": cnt["seconds_max"], "lockstep_iterations": int(cnt["iterations"]),
                "queries_solved": int(res["valid"].sum()), "host_threads": S.effective_cpus(),
                "phase_seconds": {"pop": cnt["t_pop_max"], "device+pcie": cnt["t_device_max"], "relax": cnt["t_relax_max"]},
                "release_seconds_not_in_value": cnt["t_release_max"],
                "what": "MultiQueryPlanner::plan wall time (device expansion + PCIe + host A* bookkeeping), max over ranks"}
        sys.path.insert(0, str(ROOT / "tests"))
        import planner_bindings as pb

        if pb.ref_planner_available() and args.ref_queries > 0:
            n = min(args.ref_queries, args.queries)
            t0 = time.perf_counter()
            with ThreadPoolExecutor(S.effective_cpus()) as ex:  # ctypes releases the GIL: one reference planner per thread
                outs = list(ex.map(lambda k: pb.plan_reference(make(q["start"]["pos"][k], q["goal"]["pos"][k])), range(n)))
            dt = time.perf_counter() - t0
            exp = sum(o["n_closed"] for o in outs)
            same = all(o["n_closed"] == res["n_closed"][k] and o["valid"] == res["valid"][k] for k, o in enumerate(outs))
            nt = S.effective_cpus()
            plan_s = sum(o["seconds"] for o in outs)  # inside MapPlanner::plan only (no map set-up, no teardown)
            line["reference"] = {"value": exp / (plan_s / nt), "unit": "expansions/s", "queries": n, "seconds": dt, "threads": nt,
                                 "value_incl_setup_teardown": exp / dt, "plan_seconds_sum": plan_s,
                                 "what": "the reference's MapPlanner::plan (oracle/_ref), one query per host thread",
                                 "same_results_as_gpu": bool(same)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
