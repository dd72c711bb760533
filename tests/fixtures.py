"""Shared fixtures: the corridor map (config 1) and the reference tests' parameter sets."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def corridor():
    """data/corridor.yaml via test/read_map.hpp:42-45 (data>0 -> 100 else 0)."""
    z = np.load(GOLDEN / "corridor.npz")
    grid = np.where(z["data"] > 0, 100, 0).astype(np.int8)
    return dict(start=z["start"], goal=z["goal"], origin=z["origin"], dim=z["dim"], res=float(z["resolution"]),
                grid=grid, raw=z["data"])


def U_2d(u=0.5, du=0.5):
    """test/test_planner_2d.cpp:49-53: for dx in -u..u: for dy in -u..u"""
    out = []
    dx = -u
    while dx <= u:
        dy = -u
        while dy <= u:
            out.append((dx, dy))
            dy += du
        dx += du
    return np.asarray(out, dtype=np.float64)


def U_2d_yaw(u=0.5, du=0.5, u_yaw=0.5):
    """test/test_planner_2d_with_yaw.cpp:52-57: dx, dy, dyaw nested"""
    out = []
    dx = -u
    while dx <= u:
        dy = -u
        while dy <= u:
            dyaw = -u_yaw
            while dyaw <= u_yaw:
                out.append((dx, dy, dyaw))
                dyaw += u_yaw
            dy += du
        dx += du
    return np.asarray(out, dtype=np.float64)
