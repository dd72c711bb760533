"""Benchmark / test INPUT generators (not part of the product package): deterministic synthetic inputs for
the configurations of BASELINE.json / SURVEY.md §8d.

Maps are random axis-aligned boxes + a 1-voxel boundary shell (occupied = 100, free = 0,
no unknowns), seed 42.  Control sets follow the nested-loop order of the reference tests
(test/test_planner_2d.cpp:52-53: outer loop = first axis).  Frontier nodes are reachable
lattice states: a random free cell centre propagated through a few random controls with the
reference's polynomial (include/mpl_basis/primitive.h:128-145) — inputs only.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from motion_primitive_library_b200.abi import ACC, ACCxYAW, JRK, WAYPOINT_DTYPE


def effective_cpus() -> int:
    """CPUs this process may use: os.cpu_count() capped by the cgroup v2 CPU quota (cpu.max)."""
    import os

    n = os.cpu_count() or 1
    try:
        q, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, -(-int(q) // int(period))))
    except (OSError, ValueError):
        pass
    return n


def control_set(u_max: float, n_disc: int, dim: int, yaw_rates=None) -> np.ndarray:
    """U = {-u_max..u_max}^dim (x yaw_rates), first axis outermost (reference test loop order)."""
    vals = np.linspace(-u_max, u_max, n_disc) if n_disc > 1 else np.array([0.0])
    vals = vals + 0.0  # -0.0 -> +0.0
    axes = [vals] * dim + ([np.asarray(yaw_rates, dtype=np.float64)] if yaw_rates is not None else [])
    grids = np.meshgrid(*axes, indexing="ij")
    return np.ascontiguousarray(np.stack([g.reshape(-1) for g in grids], axis=1))


def box_map(dim_cells, res, origin, n_boxes, edge_m, seed=42) -> np.ndarray:
    """x-fastest int8 grid with n_boxes random boxes of edge in [edge_m[0], edge_m[1]] metres."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = (list(dim_cells) + [1])[:3]
    three_d = len(dim_cells) == 3
    grid = np.zeros((nz, ny, nx), dtype=np.int8)  # index [z][y][x] == x + nx*y + nx*ny*z
    lo = np.asarray(origin, dtype=np.float64)
    ext = np.asarray(dim_cells, dtype=np.float64) * res
    for _ in range(n_boxes):
        c = lo + rng.random(len(dim_cells)) * ext
        e = edge_m[0] + rng.random(len(dim_cells)) * (edge_m[1] - edge_m[0])
        a = np.clip(np.floor((c - e / 2 - lo) / res).astype(int), 0, None)
        b = np.clip(np.ceil((c + e / 2 - lo) / res).astype(int), None, np.asarray(dim_cells))
        if three_d:
            grid[a[2]:b[2], a[1]:b[1], a[0]:b[0]] = 100
        else:
            grid[0, a[1]:b[1], a[0]:b[0]] = 100
    # boundary shell
    grid[:, :, 0] = grid[:, :, -1] = 100
    grid[:, 0, :] = grid[:, -1, :] = 100
    if three_d:
        grid[0, :, :] = grid[-1, :, :] = 100
    return grid.reshape(-1)


def potential_from_map_stencil(grid, dim_cells, res, radius_m, pow_=1.0) -> np.ndarray:
    """MapPlanner::createMask + updatePotentialMap on the whole map (reference
    src/mpl_planner/map_planner.cpp:286-391, 3-D branch), literally: stencil value
    (int8)(100 * ((1 - hypot(dx,dy)/rn) * (1 - |dz|/hn))^pow) for hypot<=rn, kept when > 1e-3;
    every cell with map>0 becomes 100 and stamps max() of the stencil around it.
    O(stencil x voxels): only for small maps (it pins potential_from_map in the tests)."""
    nx, ny, nz = dim_cells
    g = grid.reshape(nz, ny, nx)
    occ = g > 0
    out = g.copy()
    out[occ] = 100
    rn = int(np.ceil(radius_m[0] / res))
    hn = int(np.ceil(radius_m[2] / res))
    for dx in range(-rn, rn + 1):
        for dy in range(-rn, rn + 1):
            hyp = float(np.hypot(dx, dy))
            if hyp > rn:
                continue
            for dz in range(-hn, hn + 1):
                h = 100.0 * ((1.0 - hyp / rn) * (1.0 - abs(dz) / hn)) ** pow_
                if not h > 1e-3:
                    continue
                val = np.int8(int(h))
                if val <= 0:
                    continue
                src = occ[max(0, -dz):nz - max(0, dz), max(0, -dy):ny - max(0, dy), max(0, -dx):nx - max(0, dx)]
                dst = out[max(0, dz):nz - max(0, -dz), max(0, dy):ny - max(0, -dy), max(0, dx):nx - max(0, -dx)]
                np.maximum(dst, np.where(src, val, np.int8(0)), out=dst)
    return out.reshape(-1)


def potential_from_map(grid, dim_cells, res, radius_m, pow_=1.0) -> np.ndarray:
    """Same field as potential_from_map_stencil, computed layer-wise: the stencil value is monotone
    in (1 - hypot/rn), so for a fixed dz the maximum over occupied cells of a layer is attained at
    the nearest one — an exact 2-D Euclidean distance transform per z-layer (scipy), then a max over
    the 2*hn+1 neighbouring layers.  Input data for cfg4 (a 512^3 field takes about a minute)."""
    from scipy import ndimage

    nx, ny, nz = dim_cells
    g = grid.reshape(nz, ny, nx)
    occ = g > 0
    out = g.copy()
    out[occ] = 100
    rn = int(np.ceil(radius_m[0] / res))
    hn = int(np.ceil(radius_m[2] / res))
    d2 = np.full((nz, ny, nx), -1, dtype=np.int32)  # squared xy-distance (cells) to the layer's nearest occupied cell
    for z in range(nz):
        if occ[z].any():
            d = ndimage.distance_transform_edt(~occ[z])
            q = np.rint(d * d).astype(np.int32)
            q[d > rn] = -1
            d2[z] = q
    for dz in range(-hn, hn + 1):
        b = 1.0 - abs(dz) / hn
        src = d2[max(0, -dz):nz - max(0, dz)]
        dst = out[max(0, dz):nz - max(0, -dz)]
        hyp = np.sqrt(np.maximum(src, 0).astype(np.float64))  # == hypot(dx, dy) of the nearest occupied cell
        h = 100.0 * ((1.0 - hyp / rn) * b) ** pow_
        val = np.where((src >= 0) & (h > 1e-3), np.floor(h), 0).astype(np.int8)
        np.maximum(dst, val, out=dst)
    return out.reshape(-1)


@dataclass
class Scenario:
    name: str
    dim_cells: tuple
    res: float
    origin: tuple
    control: int
    U: np.ndarray
    T: float = 1.0
    w: float = 10.0
    wyaw: float = 1.0
    v_max: float = -1.0
    a_max: float = -1.0
    j_max: float = -1.0
    yaw_max: float = -1.0
    n_boxes: int = 0
    edge_m: tuple = (1.0, 4.0)
    potential_radius: tuple | None = None
    potential_weight: float = 0.1
    gradient_weight: float = 0.0
    seed: int = 42
    _grid: np.ndarray | None = field(default=None, repr=False)
    _pot: np.ndarray | None = field(default=None, repr=False)

    @property
    def Dim(self):
        return len(self.dim_cells)

    @property
    def nU(self):
        return int(self.U.shape[0])

    def grid(self) -> np.ndarray:
        if self._grid is None:
            self._grid = box_map(self.dim_cells, self.res, self.origin, self.n_boxes, self.edge_m, self.seed)
        return self._grid

    def potential(self):
        if self.potential_radius is None:
            return None
        if self._pot is None:
            self._pot = potential_from_map(self.grid(), self.dim_cells, self.res, self.potential_radius)
        return self._pot

    def frontier(self, n: int, seed: int = 7, max_steps: int = 6) -> np.ndarray:
        return make_frontier(self, n, seed, max_steps)


def scaled(sc: Scenario, cells: int) -> Scenario:
    """Same scenario on a smaller cube (for the CPU-checked parity tests)."""
    f = cells / sc.dim_cells[0]
    return Scenario(
        name=f"{sc.name}@{cells}", dim_cells=(cells,) * sc.Dim, res=sc.res,
        origin=tuple(-cells * sc.res / 2 for _ in range(sc.Dim)), control=sc.control, U=sc.U, T=sc.T, w=sc.w,
        wyaw=sc.wyaw, v_max=sc.v_max, a_max=sc.a_max, j_max=sc.j_max, yaw_max=sc.yaw_max,
        n_boxes=max(1, int(sc.n_boxes * f ** sc.Dim)), edge_m=sc.edge_m, potential_radius=sc.potential_radius,
        potential_weight=sc.potential_weight, gradient_weight=sc.gradient_weight, seed=sc.seed,
    )


def cfg2() -> Scenario:  # BASELINE.json configs[1]
    return Scenario("cfg2_256c_acc27", (256, 256, 256), 0.25, (-32.0, -32.0, -32.0), ACC,
                    control_set(1.0, 3, 3), v_max=3.0, n_boxes=400, edge_m=(1.0, 4.0))


def cfg3() -> Scenario:  # BASELINE.json configs[2]
    return Scenario("cfg3_512c_jrk125", (512, 512, 512), 0.1, (-25.6, -25.6, -25.6), JRK,
                    control_set(2.0, 5, 3), v_max=3.0, a_max=2.0, n_boxes=1500, edge_m=(0.5, 3.0))


def cfg_headline() -> Scenario:  # north_star target: 512^3 voxel map, 3D ACC, 27 primitives/node
    return Scenario("512c_acc27", (512, 512, 512), 0.1, (-25.6, -25.6, -25.6), ACC,
                    control_set(1.0, 3, 3), v_max=3.0, n_boxes=1500, edge_m=(0.5, 3.0))


def cfg4() -> Scenario:  # BASELINE.json configs[3]: distance-map planner, ACC x YAW, 81 primitives
    return Scenario("cfg4_512c_accyaw81_pot", (512, 512, 512), 0.1, (-25.6, -25.6, -25.6), ACCxYAW,
                    control_set(1.0, 3, 3, yaw_rates=(-0.5, 0.0, 0.5)), v_max=3.0, yaw_max=0.7, wyaw=1.0,
                    n_boxes=1500, edge_m=(0.5, 3.0), potential_radius=(1.0, 1.0, 1.0), potential_weight=0.5,
                    gradient_weight=0.0)


WORKLOADS = {"512c_acc27": cfg_headline, "cfg2": cfg2, "cfg3": cfg3, "cfg4": cfg4}


def _propagate(state, u, T, order):
    """End state of the state+control primitive (primitive.h:128-145 evaluated at T)."""
    p, v, a = state
    t2, t3 = T * T, (T * T) * T
    if order == 2:  # ACC: c3=u c4=v c5=p
        return (u / 2 * T * T + v * T + p, u * T + v, np.zeros_like(a))
    # JRK: c2=u c3=a c4=v c5=p
    return (u / 6 * t3 + a / 2 * T * T + v * T + p, u / 2 * T * T + a * T + v, u * T + a)


def make_frontier(sc: Scenario, n: int, seed: int = 7, max_steps: int = 6) -> np.ndarray:
    rng = np.random.default_rng(seed)
    D = sc.Dim
    grid = sc.grid()
    dims = np.asarray(sc.dim_cells, dtype=np.int64)
    org = np.asarray(sc.origin, dtype=np.float64)
    order = 2 if (sc.control & 15) == ACC else 3
    out = np.zeros(0, dtype=WAYPOINT_DTYPE)
    Uxyz = sc.U[:, :D]
    strides = np.array([1, dims[0], dims[0] * dims[1]][:D], dtype=np.int64)
    while out.size < n:
        m = int((n - out.size) * 1.6) + 64
        cells = (rng.random((m, D)) * dims).astype(np.int64)
        free = grid[(cells * strides).sum(1)] == 0
        cells = cells[free]
        m = cells.shape[0]
        p = (cells + 0.5) * sc.res + org
        v = np.zeros((m, D))
        a = np.zeros((m, D))
        yaw = np.zeros(m)
        steps = rng.integers(0, max_steps + 1, size=m)
        done = np.zeros(m, dtype=np.int64)
        for s in range(max_steps):
            ui = rng.integers(0, sc.nU, size=m)
            u = Uxyz[ui]
            pn, vn, an = _propagate((p, v, a), u, sc.T, order)
            ok = s < steps
            if sc.v_max > 0:
                ok &= np.abs(vn).max(1) <= sc.v_max
            if sc.a_max > 0 and order == 3:
                ok &= np.abs(an).max(1) <= sc.a_max
            cn = np.floor((pn - org) / sc.res).astype(np.int64)
            inside = ((cn >= 0) & (cn < dims)).all(1)
            ok &= inside
            ok[ok] &= grid[(cn[ok] * strides).sum(1)] == 0
            p = np.where(ok[:, None], pn, p)
            v = np.where(ok[:, None], vn, v)
            a = np.where(ok[:, None], an, a)
            if sc.U.shape[1] > D:
                yaw = np.where(ok, yaw + sc.U[ui, D] * sc.T, yaw)
            done += ok
        w = np.zeros(m, dtype=WAYPOINT_DTYPE)
        w["pos"][:, :D] = p
        w["vel"][:, :D] = v
        w["acc"][:, :D] = a
        w["yaw"] = np.arctan2(np.sin(yaw), np.cos(yaw)) if sc.U.shape[1] > D else 0.0
        w["t"] = done * sc.T
        out = np.concatenate([out, w])
    return np.ascontiguousarray(out[:n])
