"""ctypes binding of libmpl_host.so: the C++ host planner (MPL::MapPlanner with the GPU env,
MPL::MultiQueryPlanner) behind flat C structs.  See host/plan_capi.hpp."""
import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB = PKG / "lib" / "libmpl_host.so"

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
        ("gradient_weight", C.c_double), ("heur_ignore_dynamics", C.c_int32),
    ]


class PlanResult(C.Structure):
    _fields_ = [("valid", C.c_int32), ("cost", C.c_double), ("expanded", C.c_int32), ("n_closed", C.c_int32),
                ("n_open", C.c_int32), ("n_actions", C.c_int32), ("gpu_nodes", C.c_int64), ("gpu_calls", C.c_int64),
                ("gpu_launches", C.c_int64), ("seconds", C.c_double)]


class QueryResult(C.Structure):
    _fields_ = [("valid", C.c_int32), ("cost", C.c_double), ("expanded", C.c_int32), ("n_closed", C.c_int32),
                ("n_actions", C.c_int32)]


class LpaStep(C.Structure):
    _fields_ = [("op", C.c_int32), ("n", C.c_int32), ("cells", C.c_void_p)]


class LpaOut(C.Structure):
    _fields_ = [("valid", C.c_int32), ("cost", C.c_double), ("expanded", C.c_int32), ("n_actions", C.c_int32),
                ("n_states", C.c_int32), ("n_closed", C.c_int32), ("n_open", C.c_int32), ("state_hash", C.c_uint64),
                ("n_linked", C.c_int64), ("linked_hash", C.c_uint64), ("seconds", C.c_double)]


OP_PLAN, OP_LINK, OP_BLOCK, OP_CLEAR, OP_SUBTREE = range(5)

WAYPOINT_DTYPE = np.dtype(
    [("pos", "<f8", 3), ("vel", "<f8", 3), ("acc", "<f8", 3), ("jrk", "<f8", 3), ("yaw", "<f8"), ("t", "<f8")]
)


def load_fn(path, fn):
    L = C.CDLL(str(path))
    f = getattr(L, fn)
    f.argtypes = [C.POINTER(PlanArgs), C.POINTER(PlanResult), C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    f.restype = C.c_int
    return L, f


def make_args(dim, control, grid, mdim, origin, res, U, start, goal, T=1.0, w=10.0, wyaw=1.0, eps=1.0, v_max=-1.0,
              a_max=-1.0, j_max=-1.0, yaw_max=-1.0, tol_pos=0.5, tol_vel=-1.0, tol_acc=-1.0, max_num=-1, speculate=0,
              potential=None, potential_weight=0.1, gradient_weight=0.0, heur_ignore_dynamics=True):
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
    a.heur_ignore_dynamics = 1 if heur_ignore_dynamics else 0
    a._keep = keep
    return a


def run_plan(fn, lib, args, cap=1 << 21):
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


def run_lpa(fn, lib, args, script, cap_actions=4096):
    """Run a scripted LPA* session.  script: list of ("plan",), ("link",), ("block", cells[n, dim]),
    ("clear", cells[n, dim]), ("subtree", k).  Returns one dict per step (fields of mplh_lpa_out, plus
    the action ids of the trajectory for plan steps)."""
    ops = dict(plan=OP_PLAN, link=OP_LINK, block=OP_BLOCK, clear=OP_CLEAR, subtree=OP_SUBTREE)
    steps = (LpaStep * len(script))()
    keep = []
    for k, st in enumerate(script):
        steps[k].op = ops[st[0]]
        if st[0] in ("block", "clear"):
            cells = np.ascontiguousarray(st[1], dtype=np.int32).reshape(-1, args.dim)
            keep.append(cells)
            steps[k].n = len(cells)
            steps[k].cells = cells.ctypes.data
        elif st[0] == "subtree":
            steps[k].n = int(st[1])
    outs = (LpaOut * len(script))()
    actions = np.full((len(script), cap_actions), -1, dtype=np.int32)
    rc = fn(C.byref(args), steps, len(script), outs, actions.ctypes.data, cap_actions)
    if rc != 0:
        err = getattr(lib, "mplh_last_error", None)
        raise RuntimeError(err().decode() if err else f"lpa session failed rc={rc}")
    res = []
    for k in range(len(script)):
        d = {name: getattr(outs[k], name) for name, _ in LpaOut._fields_}
        d["op"] = script[k][0]
        d["actions"] = actions[k, : min(d["n_actions"], cap_actions)].copy()
        res.append(d)
    return res


def load_lpa_fn(path, fn):
    L = C.CDLL(str(path))
    f = getattr(L, fn)
    f.argtypes = [C.POINTER(PlanArgs), C.POINTER(LpaStep), C.c_int, C.POINTER(LpaOut), C.c_void_p, C.c_int]
    f.restype = C.c_int
    return L, f


def lpa_session(args, script):
    """MPL::MapPlanner with setLPAstar(true) and the GPU env: plan / getLinkedNodes / updateBlockedNodes /
    updateClearedNodes / getSubStateSpace as scripted; expansion, the linked-voxel walk and the
    is_free(pr) re-validation run on the device."""
    if not LIB.exists():
        raise ImportError(f"{LIB} not built (python -c 'import __graft_entry__ as g; g.build()')")
    lib, fn = load_lpa_fn(LIB, "mplh_lpa_run")
    lib.mplh_last_error.restype = C.c_char_p
    return run_lpa(fn, lib, args, script)


def load_traj_fn(path, fn):
    L = C.CDLL(str(path))
    f = getattr(L, fn)
    f.argtypes = [C.POINTER(PlanArgs), C.c_int, C.POINTER(PlanResult), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                  C.POINTER(C.c_int32), C.c_void_p]
    f.restype = C.c_int
    return L, f


def run_trajectory(fn, lib, args, n_samples=50, cap_wp=4096):
    """plan(), then the recovered Trajectory (trajectory.h): dict with `commands` = sample(N) rows
    {pos, vel, acc, jrk, yaw, yaw_dot, t}, `total_time`, `J`, `Jyaw`, `segments`, `waypoints` =
    getWaypoints() rows {pos, vel, acc, jrk, yaw, t} and `evaluated` = evaluate(t) at the sample times."""
    dim = args.dim
    r = PlanResult()
    samples = np.zeros((n_samples + 1, 4 * dim + 3))
    totals = np.zeros(4)
    wps = np.zeros((cap_wp, 4 * dim + 2))
    mids = np.zeros((n_samples + 1, 4 * dim + 2))
    n_wp = C.c_int32(0)
    rc = fn(C.byref(args), n_samples, C.byref(r), samples.ctypes.data, totals.ctypes.data, wps.ctypes.data, cap_wp,
            C.byref(n_wp), mids.ctypes.data)
    if rc != 0:
        err = getattr(lib, "mplh_last_error", None)
        raise RuntimeError(err().decode() if err else f"trajectory failed rc={rc}")
    out = {k: getattr(r, k) for k, _ in PlanResult._fields_}
    out.update(commands=samples, total_time=totals[0], J=totals[1], Jyaw=totals[2], segments=int(totals[3]),
               waypoints=wps[: min(n_wp.value, cap_wp)].copy(), evaluated=mids)
    return out


def plan_trajectory(args, n_samples=50):
    """MPL::MapPlanner::plan() with the GPU env + the recovered Trajectory (sample / waypoints / efforts)."""
    if not LIB.exists():
        raise ImportError(f"{LIB} not built (python -c 'import __graft_entry__ as g; g.build()')")
    lib, fn = load_traj_fn(LIB, "mplh_plan_trajectory")
    lib.mplh_last_error.restype = C.c_char_p
    return run_trajectory(fn, lib, args, n_samples)


def load_iter_fn(path, fn):
    L = C.CDLL(str(path))
    f = getattr(L, fn)
    f.argtypes = [C.POINTER(PlanArgs), C.c_void_p, C.c_int, C.POINTER(PlanResult), C.POINTER(PlanResult), C.c_void_p,
                  C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    f.restype = C.c_int
    return L, f


def run_iterative(fn, lib, args, search_radius, max_iter=3, cap=1 << 21):
    """plan() followed by MapPlanner::iterativePlan() from that trajectory.  Returns (first, last):
    dicts of the plan-result fields; `last` also has closed keys, action ids, `iterations` (plan() calls
    iterativePlan made; 0 when the driver cannot know) and `ok` (its return value)."""
    first, last = PlanResult(), PlanResult()
    info = np.zeros(2, dtype=np.int32)
    rad = np.zeros(3)
    rad[: len(search_radius)] = search_radius
    closed = np.zeros(cap, dtype=np.uint64)
    actions = np.zeros(65536, dtype=np.int32)
    rc = fn(C.byref(args), rad.ctypes.data, int(max_iter), C.byref(first), C.byref(last), info.ctypes.data,
            closed.ctypes.data, cap, actions.ctypes.data, actions.size)
    if rc != 0:
        err = getattr(lib, "mplh_last_error", None)
        raise RuntimeError(err().decode() if err else f"iterative plan failed rc={rc}")
    f = {k: getattr(first, k) for k, _ in PlanResult._fields_}
    l = {k: getattr(last, k) for k, _ in PlanResult._fields_}
    l["closed"] = closed[: min(last.n_closed, cap)].copy()
    l["actions"] = actions[: last.n_actions].copy()
    l["iterations"], l["ok"] = int(info[0]), int(info[1])
    return f, l


def iterative_plan(args, search_radius, max_iter=3):
    """MPL::MapPlanner::plan() + iterativePlan() with the GPU env (tunnels built on the device)."""
    if not LIB.exists():
        raise ImportError(f"{LIB} not built (python -c 'import __graft_entry__ as g; g.build()')")
    lib, fn = load_iter_fn(LIB, "mplh_iterative_plan")
    lib.mplh_last_error.restype = C.c_char_p
    return run_iterative(fn, lib, args, search_radius, max_iter)


def _host():
    if not LIB.exists():
        raise ImportError(f"{LIB} not built (python -c 'import __graft_entry__ as g; g.build()')")
    lib, fn = load_fn(LIB, "mplh_plan")
    lib.mplh_last_error.restype = C.c_char_p
    lib.mplh_plan_batch.argtypes = [C.POINTER(PlanArgs), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.mplh_plan_batch.restype = C.c_int
    lib.mplh_batch_open.argtypes = [C.POINTER(PlanArgs)]
    lib.mplh_batch_open.restype = C.c_void_p
    lib.mplh_batch_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    lib.mplh_batch_plan.restype = C.c_int
    lib.mplh_batch_close.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.mplh_batch_close.restype = C.c_int
    return lib, fn


def plan(args):
    """MPL::MapPlanner<Dim>::plan() with the GPU env (one query)."""
    lib, fn = _host()
    return run_plan(fn, lib, args)


def plan_trace(args, cap=1 << 20):
    """plan() with the GPU env, returning also the expanded nodes in A* pop order (WAYPOINT_DTYPE array)."""
    lib, _ = _host()
    lib.mplh_plan_trace.argtypes = [C.POINTER(PlanArgs), C.POINTER(PlanResult), C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
    lib.mplh_plan_trace.restype = C.c_int
    r = PlanResult()
    trace = np.zeros(cap, dtype=WAYPOINT_DTYPE)
    n = C.c_int32(0)
    rc = lib.mplh_plan_trace(C.byref(args), C.byref(r), trace.ctypes.data, cap, C.byref(n))
    if rc != 0:
        raise RuntimeError(lib.mplh_last_error().decode())
    out = {k: getattr(r, k) for k, _ in PlanResult._fields_}
    out["trace"] = trace[: n.value].copy()
    return out


def plan_batch(args, starts, goals):
    """MPL::MultiQueryPlanner: lock-step A* over many (start, goal) pairs; one device launch per
    iteration expands the current node of every live query.  starts/goals: WAYPOINT_DTYPE arrays."""
    lib, _ = _host()
    starts = np.ascontiguousarray(starts, dtype=WAYPOINT_DTYPE)
    goals = np.ascontiguousarray(goals, dtype=WAYPOINT_DTYPE)
    nq = len(starts)
    out = (QueryResult * max(nq, 1))()
    totals = np.zeros(7)
    rc = lib.mplh_plan_batch(C.byref(args), starts.ctypes.data, goals.ctypes.data, nq, out, totals.ctypes.data)
    if rc != 0:
        raise RuntimeError(lib.mplh_last_error().decode())
    res = np.zeros(nq, dtype=[("valid", "i4"), ("cost", "f8"), ("expanded", "i4"), ("n_closed", "i4"), ("n_actions", "i4")])
    for q in range(nq):
        res[q] = (out[q].valid, out[q].cost, out[q].expanded, out[q].n_closed, out[q].n_actions)
    return res, dict(iterations=int(totals[0]), nodes=int(totals[1]), seconds=float(totals[2]), t_pop=float(totals[3]),
                     t_device=float(totals[4]), t_relax=float(totals[5]), t_release=float(totals[6]))


_QRES = [("valid", "i4"), ("cost", "f8"), ("expanded", "i4"), ("n_closed", "i4"), ("n_actions", "i4")]


class BatchPlanner:
    """A MPL::MultiQueryPlanner session: the map is uploaded once, and the search states of one query set
    are recycled for the next (a planner that answers batch after batch allocates its state memory once).
    `args` supplies the map, the controls and the limits (its start/goal are ignored)."""

    def __init__(self, args):
        self._lib, _ = _host()
        self._args = args  # keeps the arrays the struct points at alive
        self._h = self._lib.mplh_batch_open(C.byref(args))
        if not self._h:
            raise RuntimeError(self._lib.mplh_last_error().decode())

    def plan(self, starts, goals, eps=None, max_num=None):
        starts = np.ascontiguousarray(starts, dtype=WAYPOINT_DTYPE)
        goals = np.ascontiguousarray(goals, dtype=WAYPOINT_DTYPE)
        nq = len(starts)
        out = (QueryResult * max(nq, 1))()
        totals = np.zeros(7)
        rc = self._lib.mplh_batch_plan(self._h, starts.ctypes.data, goals.ctypes.data, nq,
                                       self._args.eps if eps is None else eps,
                                       self._args.max_num if max_num is None else max_num, out, totals.ctypes.data)
        if rc != 0:
            raise RuntimeError(self._lib.mplh_last_error().decode())
        res = np.zeros(nq, dtype=_QRES)
        for q in range(nq):
            res[q] = (out[q].valid, out[q].cost, out[q].expanded, out[q].n_closed, out[q].n_actions)
        return res, dict(iterations=int(totals[0]), nodes=int(totals[1]), seconds=float(totals[2]), t_pop=float(totals[3]),
                         t_device=float(totals[4]), t_relax=float(totals[5]), t_release=0.0)

    def close(self) -> float:
        """Free the session (incl. the kept search states); returns the seconds that took."""
        if not self._h:
            return 0.0
        rel = C.c_double(0.0)
        self._lib.mplh_batch_close(self._h, C.byref(rel))
        self._h = None
        return rel.value

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
