#!/usr/bin/env python
"""Wall time of one MPL::MapPlanner::plan() (host A* + env) on a synthetic 3-D voxel map:
the CPU checker env (oracle) vs the GPU env with different speculation depths.
Usage: python tools/plan_timing.py [cells] [max_expansions]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import planner_bindings as pb  # noqa: E402
import scenarios as S  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 256
maxn = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
for name, sc in (("ACC-27", S.scaled(S.cfg_headline(), cells)), ("JRK-125", S.scaled(S.cfg3(), cells))):
    grid = sc.grid()
    nodes = sc.frontier(256, seed=12, max_steps=0)
    d = np.abs(nodes["pos"][:, None, :] - nodes["pos"][None, :, :]).max(-1)
    i, j = np.unravel_index(np.argmax(d), d.shape)

    def args(k):
        return pb.make_args(3, sc.control, grid, sc.dim_cells, sc.origin, sc.res, sc.U, start=dict(pos=nodes["pos"][i]),
                            goal=dict(pos=nodes["pos"][j]), v_max=sc.v_max, a_max=sc.a_max, T=sc.T, w=sc.w, speculate=k,
                            max_num=maxn)

    ref = pb.plan_oracle(args(1))
    print(f"{name} {cells}^3: CPU env   expanded {ref['expanded']:6d} valid {ref['valid']} cost {ref['cost']:.3f} "
          f"time {ref['seconds']*1e3:8.1f} ms  ({ref['expanded']/ref['seconds']/1e3:.1f} k exp/s)")
    for k in (1, 16, 128, 1024):
        g = pb.plan_gpu(args(k))
        same = g["expanded"] == ref["expanded"] and np.array_equal(g["closed"], ref["closed"])
        print(f"{name} {cells}^3: GPU env K={k:4d} expanded {g['expanded']:6d} launches {g['gpu_calls']:6d} nodes sent "
              f"{g['gpu_nodes']:7d} time {g['seconds']*1e3:8.1f} ms  ({g['expanded']/g['seconds']/1e3:.1f} k exp/s) same_closed_set={same}")
