#!/usr/bin/env python
"""Per-step wall time of a scripted LPA* replanning session (plan, getLinkedNodes, block a blob on
the trajectory, replan, clear it, replan) on a synthetic 3-D voxel map: the reference's own planner
(oracle/_ref, CPU, when present) vs this repository's host planner with the GPU env.
Usage: python tools/lpa_timing.py [cells] [max_expansions] [speculate]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import planner_bindings as pb  # noqa: E402
import scenarios as S  # noqa: E402
from test_lpastar_vs_ref import FIELDS, integrate_cells  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 256
maxn = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
spec = int(sys.argv[3]) if len(sys.argv) > 3 else 16
for name, sc in (("ACC-27", S.scaled(S.cfg_headline(), cells)), ("JRK-125", S.scaled(S.cfg3(), cells))):
    grid = sc.grid()
    nodes = sc.frontier(256, seed=12, max_steps=0)
    d = np.abs(nodes["pos"][:, None, :] - nodes["pos"][None, :, :]).max(-1)
    i, j = np.unravel_index(np.argmax(d), d.shape)
    a = pb.make_args(3, sc.control, grid, sc.dim_cells, sc.origin, sc.res, sc.U, start=dict(pos=nodes["pos"][i]),
                     goal=dict(pos=nodes["pos"][j]), v_max=sc.v_max, a_max=sc.a_max, T=sc.T, w=sc.w, speculate=spec,
                     max_num=maxn)
    first = pb.lpa_session(a, [("plan",)])[0]
    if not first["valid"]:
        print(f"{name} {cells}^3: no trajectory within {maxn} expansions")
        continue
    tc = integrate_cells(sc.control, sc.U, nodes["pos"][i], first["actions"], sc.origin, sc.res)
    mid = tc[len(tc) // 2]
    blob = np.array([mid + (dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)], dtype=np.int32)
    script = [("plan",), ("link",), ("block", blob), ("plan",), ("link",), ("clear", blob), ("plan",)]
    g = pb.lpa_session(a, script)
    r = pb.lpa_reference(a, script) if pb.ref_planner_available() else None
    same = r is not None and all(x[f] == y[f] for x, y in zip(g, r) for f in FIELDS)
    print(f"{name} {cells}^3 (speculate {spec}): states {g[0]['n_states']}, stored edges walked by LINK -> {g[1]['n_linked']} voxels"
          + (f"; identical search state after every step: {same}" if r else ""))
    for k, st in enumerate(script):
        line = f"  {st[0]:6s} gpu-env {g[k]['seconds']*1e3:9.2f} ms"
        if r:
            line += f"   reference {r[k]['seconds']*1e3:9.2f} ms   x{r[k]['seconds']/max(g[k]['seconds'],1e-9):6.1f}"
        if st[0] == "plan":
            line += f"   valid {g[k]['valid']} cost {g[k]['cost']:.3f}"
        print(line)
