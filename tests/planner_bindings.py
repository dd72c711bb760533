"""ctypes binding of the one-shot planner entry points:
   mplh_plan  (motion_primitive_library_b200/lib/libmpl_host.so — product: host A* + GPU env)
   orcp_plan  (oracle/liboracle_planner.so — TEST: same host A*, CPU-oracle env)"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


class Waypoint(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("vel", C.c_double * 3), ("acc", C.c_double * 3), ("jrk", C.c_double * 3),
                ("yaw", C.c_double), ("t", C.c_double)]


class PlanArgs(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("control", C.c_int32), ("map", C.c_void_p), ("mdim", C.c_int32 * 3),
        ("origin", C.c_double * 3), ("res", C.c_double), ("U", C.c_void_p), ("nU", C.c_int32), ("udim", C.c_int32),
        ("T", C.c_double), ("w", C.c_double), ("wyaw", C.c_double), ("eps", C.c_double),
        ("v_max", C.c_double), ("a_max", C.c_double), ("j_max", C.c_double), ("yaw_max", C.c_double),
        ("tol_pos", C.c_double), ("tol_vel", C.c_double), ("tol_acc", C.c_double),
        ("start", Waypoint), ("goal", Waypoint), ("max_num", C.c_int32), ("speculate", C.c_int32),
        ("device", C.c_int32), ("potential", C.c_void_p), ("potential_weight", C.c_double),
        ("gradient_weight", C.c_double),
    ]


class PlanResult(C.Structure):
    _fields_ = [("valid", C.c_int32), ("cost", C.c_double), ("expanded", C.c_int32), ("n_closed", C.c_int32),
                ("n_open", C.c_int32), ("n_actions", C.c_int32), ("gpu_nodes", C.c_int64), ("gpu_calls", C.c_int64),
                ("gpu_launches", C.c_int64), ("seconds", C.c_double)]


def _load(path, fn):
    L = C.CDLL(str(path))
    f = getattr(L, fn)
    f.argtypes = [C.POINTER(PlanArgs), C.POINTER(PlanResult), C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    f.restype = C.c_int
    return L, f


def make_args(dim, control, grid, mdim, origin, res, U, start, goal, T=1.0, w=10.0, wyaw=1.0, eps=1.0, v_max=-1.0,
              a_max=-1.0, j_max=-1.0, yaw_max=-1.0, tol_pos=0.5, tol_vel=-1.0, tol_acc=-1.0, max_num=-1, speculate=1,
              potential=None, potential_weight=0.1, gradient_weight=0.0):
    keep = dict(grid=np.ascontiguousarray(grid, dtype=np.int8), U=np.ascontiguousarray(U, dtype=np.float64))
    a = PlanArgs()
    a.dim, a.control = dim, control
    a.map = keep["grid"].ctypes.data
    for k in range(3):
        a.mdim[k] = int(mdim[k]) if k < dim else 1
        a.origin[k] = float(origin[k]) if k < dim else 0.0
    a.res = res
    a.U = keep["U"].ctypes.data
    a.nU, a.udim = keep["U"].shape
    a.T, a.w, a.wyaw, a.eps = T, w, wyaw, eps
    a.v_max, a.a_max, a.j_max, a.yaw_max = v_max, a_max, j_max, yaw_max
    a.tol_pos, a.tol_vel, a.tol_acc = tol_pos, tol_vel, tol_acc
    for name, src in (("start", start), ("goal", goal)):
        w_ = getattr(a, name)
        for f in ("pos", "vel", "acc", "jrk"):
            v = src.get(f, ())
            for k in range(len(v)):
                getattr(w_, f)[k] = float(v[k])
        w_.yaw = float(src.get("yaw", 0.0))
    a.max_num, a.speculate, a.device = max_num, speculate, 0
    if potential is not None:
        keep["pot"] = np.ascontiguousarray(potential, dtype=np.int8)
        a.potential = keep["pot"].ctypes.data
    a.potential_weight, a.gradient_weight = potential_weight, gradient_weight
    a._keep = keep
    return a


def _run(fn, lib, args, cap=1 << 21):
    r = PlanResult()
    closed = np.zeros(cap, dtype=np.uint64)
    actions = np.zeros(65536, dtype=np.int32)
    rc = fn(C.byref(args), C.byref(r), closed.ctypes.data, cap, actions.ctypes.data, actions.size)
    if rc != 0:
        err = getattr(lib, "mplh_last_error", None)
        raise RuntimeError(err().decode() if err else f"plan failed rc={rc}")
    out = {k: getattr(r, k) for k, _ in PlanResult._fields_}
    out["closed"] = closed[: min(r.n_closed, cap)].copy()
    out["actions"] = actions[: r.n_actions].copy()
    return out


def plan_gpu(args):
    lib, fn = _load(ROOT / "motion_primitive_library_b200" / "lib" / "libmpl_host.so", "mplh_plan")
    lib.mplh_last_error.restype = C.c_char_p
    return _run(fn, lib, args)


def plan_oracle(args):
    p = ROOT / "oracle" / "liboracle_planner.so"
    if not p.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "liboracle_planner.so"], stdout=subprocess.DEVNULL)
    lib, fn = _load(p, "orcp_plan")
    return _run(fn, lib, args)
