#!/usr/bin/env python
"""Wall time of MapPlanner::updatePotentialMap: device (mplx_update_potential_map, incl. the D2H of
the new grid) vs the reference (oracle/_ref planner library) on the same synthetic voxel map.
Usage: python tools/maps_timing.py [cells=256] [radius_m=1.0]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import planner_bindings as pb  # noqa: E402
from motion_primitive_library_b200 import MapUtil, env_map  # noqa: E402
import scenarios as S  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rad = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
sc = S.scaled(S.cfg_headline(), cells)
g = sc.grid()
mu = MapUtil()
mu.setMap(sc.origin, sc.dim_cells, g, sc.res)
e = env_map(mu)
t0 = time.perf_counter()
out = e.update_potential_map((rad, rad, rad))
t1 = time.perf_counter()
print(f"{cells}^3 res {sc.res} radius {rad} m: occupied {(g > 0).mean():.3f}; device updatePotentialMap {1e3 * (t1 - t0):.1f} ms "
      f"(3 kernels + D2H of {g.size / 1e6:.0f} MB); cells with potential {(out > 0).mean():.3f}")
if pb.ref_planner_available() and cells <= 256:
    a = pb.make_args(3, 0x03, g, sc.dim_cells, sc.origin, sc.res, np.zeros((1, 3)), start=dict(pos=(0, 0, 0)), goal=dict(pos=(0, 0, 0)))
    t0 = time.perf_counter()
    ref = pb.reference_potential_map(a, (rad, rad, rad), g.size)
    t1 = time.perf_counter()
    print(f"reference updatePotentialMap (1 core): {1e3 * (t1 - t0):.0f} ms; identical: {np.array_equal(ref, out)}")
