"""GPU parity: libmplx (CUDA, through the C ABI with host buffers) vs the CPU oracle.

Bar: successor Waypoints, lattice ints, hash keys, action ids and counts bit-exact;
edge costs within 1e-6 relative with identical +inf pattern (north_star).
"""
import numpy as np
import pytest

import fixtures
import oracle_bindings as ob
from parity import assert_expansion_equal

pytestmark = pytest.mark.gpu

VEL, ACC, JRK, SNP, ACCxYAW, JRKxYAW = 0x01, 0x03, 0x07, 0x0F, 0x13, 0x17
WANT = ("succ", "cost", "action", "key", "lattice")


def gpu_env(sc, region=None):
    from motion_primitive_library_b200 import MapUtil, env_map

    mu = MapUtil()
    mu.setMap(sc.origin, sc.dim_cells, sc.grid(), sc.res)
    e = env_map(mu)
    e.set_control(sc.control)
    e.set_u(sc.U)
    e.set_dt(sc.T)
    e.set_w(sc.w)
    e.set_wyaw(sc.wyaw)
    e.set_v_max(sc.v_max)
    e.set_a_max(sc.a_max)
    e.set_j_max(sc.j_max)
    e.set_yaw_max(sc.yaw_max)
    if sc.potential() is not None:
        e.set_potential_weight(sc.potential_weight)
        e.set_gradient_weight(sc.gradient_weight)
        e.set_potential_map(sc.potential())
    if region is not None:
        e.set_search_region(region)
    return e


def run_parity(sc, n, region=None, exact_cost=False, seed=7, kernels=(1, 2, 3, 4, 5, 0)):
    """All kernels (1 = literal sequential loop, 2 = register kernel, 3 = flat sample-parallel, 4 = dealing,
    5 = fixed-point sample loop where the plan allows it, 0 = what the auto rule picks) vs the oracle."""
    nodes = sc.frontier(n, seed=seed)
    orc = ob.OracleEnv.from_scenario(sc, region=region).expand(nodes, nthreads=8)
    env = gpu_env(sc, region)
    for which in kernels:
        env.set_kernel(which)
        g = env.expand(nodes, want=WANT)
        st = assert_expansion_equal(g, orc, exact_cost=exact_cost)
        assert st["successors"] > 0
    return st, g, orc


def test_headline_acc27_small_map():
    import scenarios as S

    st, g, orc = run_parity(S.scaled(S.cfg_headline(), 128), 6000, exact_cost=True)
    inf_frac = np.isinf(orc["cost"][: 27]).mean()
    assert st["finite"] < st["successors"], "map has obstacles: some +inf entries expected"


def test_cfg2_acc27_coarse_map():
    import scenarios as S

    run_parity(S.scaled(S.cfg2(), 96), 4000, exact_cost=True)


def test_cfg3_jrk125():
    import scenarios as S

    run_parity(S.scaled(S.cfg3(), 128), 1500, exact_cost=True)


def test_cfg4_accyaw81_potential():
    import scenarios as S

    run_parity(S.scaled(S.cfg4(), 96), 1500)


def test_cfg4_with_gradient_weight():
    import scenarios as S

    sc = S.scaled(S.cfg4(), 64)
    sc.gradient_weight = 0.3
    run_parity(sc, 800)


def test_search_region_tunnel():
    import scenarios as S

    sc = S.scaled(S.cfg_headline(), 64)
    rng = np.random.default_rng(3)
    region = (rng.random(64 ** 3) < 0.8).astype(np.uint8)
    st, g, orc = run_parity(sc, 1500, region=region, exact_cost=True)


def _custom(dim, control, U, cells, res, seed=5, **kw):
    from scenarios import Scenario

    return Scenario(f"custom{dim}d_{control:x}", (cells,) * dim, res, tuple(-cells * res / 2 for _ in range(dim)),
                    control, U, n_boxes=max(2, cells // 6), edge_m=(4 * res, 12 * res), seed=seed, **kw)


def test_corridor_2d_acc9_reference_test_config():
    """config 1: test/test_planner_2d.cpp parameters on data/corridor.yaml."""
    from motion_primitive_library_b200 import MapUtil, env_map
    from motion_primitive_library_b200.abi import WAYPOINT_DTYPE

    c = fixtures.corridor()
    U = fixtures.U_2d()
    orc_env = ob.OracleEnv(2, ACC, U, c["grid"], c["dim"], c["origin"], c["res"], T=1.0, w=10.0, v_max=1.0, a_max=1.0)
    # frontier: BFS wavefront from the start through the oracle (finite-cost successors only)
    front = np.zeros(1, dtype=WAYPOINT_DTYPE)
    front["pos"][0, :2] = c["start"]
    allnodes = [front]
    for _ in range(6):
        r = orc_env.expand(front, nthreads=8)
        nU = r["nU"]
        valid = (np.arange(nU)[None, :] < r["count"][:, None]).reshape(-1) & ~np.isinf(r["cost"])
        nxt = r["succ"][valid]
        _, uniq = np.unique(r["key"][valid], return_index=True)
        front = nxt[np.sort(uniq)]
        allnodes.append(front)
    nodes = np.concatenate(allnodes)
    assert nodes.size > 500
    mu = MapUtil()
    mu.setMap(c["origin"], c["dim"], c["grid"], c["res"])
    e = env_map(mu)
    e.set_control(ACC)
    e.set_u(U)
    e.set_v_max(1.0)
    e.set_a_max(1.0)
    e.set_dt(1.0)
    g = e.expand(nodes, want=WANT)
    assert_expansion_equal(g, orc_env.expand(nodes, nthreads=8), exact_cost=True)
    # SURVEY.md §10 known answers, now from the GPU
    s, cst, act = g.node(0)
    assert list(act) == [0, 1, 2, 3, 5, 6, 7, 8]
    assert list(cst) == [10.5, 10.25, 10.5, 10.25, 10.25, 10.5, 10.25, 10.5]
    assert tuple(g.lattice[0][:4]) == (225, -5, -375, -5)


def test_corridor_2d_accyaw27_reference_test_config():
    """test/test_planner_2d_with_yaw.cpp parameters (yaw_max 0.7, 27 controls)."""
    from scenarios import Scenario

    c = fixtures.corridor()
    sc = Scenario("corridor_yaw", tuple(int(x) for x in c["dim"]), c["res"], tuple(c["origin"]), ACCxYAW,
                  fixtures.U_2d_yaw(), v_max=1.0, a_max=1.0, yaw_max=0.7)
    sc._grid = c["grid"]
    run_parity(sc, 3000)


def test_2d_vel_and_3d_snp_and_jrkyaw():
    from scenarios import control_set

    # frontier generator handles ACC/JRK; for VEL/SNP build nodes by hand on a free map
    rng = np.random.default_rng(11)
    for dim, control, U, kw in (
        (2, VEL, control_set(1.0, 3, 2), {}),
        (3, SNP, control_set(4.0, 3, 3), dict(v_max=2.5, a_max=3.0, j_max=6.0)),
        (3, JRKxYAW, control_set(2.0, 3, 3, yaw_rates=(-0.4, 0.0, 0.4)), dict(v_max=3.0, a_max=2.0, yaw_max=1.0)),
    ):
        sc = _custom(dim, control, U, 48, 0.2, **kw)
        n = 1200
        nodes = np.zeros(n, dtype=ob.WAYPOINT_DTYPE)
        half = 48 * 0.2 / 2
        nodes["pos"][:, :dim] = np.round(rng.uniform(-half * 0.9, half * 0.9, (n, dim)) / 0.05) * 0.05
        nodes["vel"][:, :dim] = rng.integers(-4, 5, (n, dim)) * 0.5
        nodes["acc"][:, :dim] = rng.integers(-4, 5, (n, dim)) * 0.5
        nodes["jrk"][:, :dim] = rng.integers(-3, 4, (n, dim)) * 1.0
        nodes["yaw"] = rng.integers(-7, 8, n) * 0.4
        nodes["t"] = rng.integers(0, 5, n) * 1.0
        orc = ob.OracleEnv.from_scenario(sc).expand(nodes, nthreads=8)
        env = gpu_env(sc)
        for which in (1, 2, 3, 4, 5):
            env.set_kernel(which)
            g = env.expand(nodes, want=WANT)
            st = assert_expansion_equal(g, orc, exact_cost=(control & 16) == 0)
            assert st["successors"] > n


def test_empty_and_ragged_batches_and_pinned_buffers():
    import scenarios as S

    sc = S.scaled(S.cfg_headline(), 64)
    e = gpu_env(sc)
    orc_env = ob.OracleEnv.from_scenario(sc)
    r = e.expand(np.zeros(0, dtype=ob.WAYPOINT_DTYPE))
    assert r.count.size == 0
    for which in (2, 0):
        e.set_kernel(which)
        for n in (1, 2, 9, 10, 31, 257):  # not multiples of the nodes-per-CTA (9 for |U|=27)
            nodes = sc.frontier(n, seed=n)
            assert_expansion_equal(e.expand(nodes, want=WANT), orc_env.expand(nodes), exact_cost=True)
    nodes = sc.frontier(3000, seed=1)
    assert_expansion_equal(e.expand(nodes, want=WANT, pinned=True), orc_env.expand(nodes, nthreads=8), exact_cost=True)
    # single-node get_succ (the reference signature)
    s, c, a = e.get_succ(nodes[5])
    o = orc_env.get_succ(nodes[5])
    np.testing.assert_array_equal(a, o["action"])
    assert s.tobytes() == o["succ"].tobytes()


def test_setter_invalidation_and_stats():
    import scenarios as S

    sc = S.scaled(S.cfg_headline(), 64)
    e = gpu_env(sc)
    nodes = sc.frontier(500, seed=2)
    e.enable_stats(True)
    t = ob.OracleEnv.from_scenario(sc).timed(nodes)
    for which in (2, 4, 0):
        e.set_kernel(which)
        g = e.expand(nodes, want=WANT)
        samples, succ = e.last_stats()
        assert samples == t["samples"] and succ == t["successors"] == int(g.count.sum())
    e.enable_stats(False)
    # change a limit: results must follow (params re-uploaded)
    e.set_v_max(1.0)
    sc.v_max = 1.0
    assert_expansion_equal(e.expand(nodes, want=WANT), ob.OracleEnv.from_scenario(sc).expand(nodes), exact_cost=True)
    assert e.launch_count() >= 2


def test_unbounded_velocity_falls_back_past_the_sample_table():
    """v_max <= 0 (unlimited, env_base.h:380) with fast nodes: n = ceil(max_v*T/res) exceeds the
    128-row sample-time table for some primitives -> in-kernel sequential fallback."""
    from scenarios import control_set

    sc = _custom(3, ACC, control_set(1.0, 3, 3), 64, 0.05)
    sc.v_max = -1.0
    rng = np.random.default_rng(5)
    n = 600
    nodes = np.zeros(n, dtype=ob.WAYPOINT_DTYPE)
    nodes["pos"][:, :3] = np.round(rng.uniform(-1.2, 1.2, (n, 3)) / 0.05) * 0.05
    nodes["vel"][:, :3] = rng.integers(-9, 10, (n, 3)) * 1.0  # up to 9 m/s / 0.05 m = n up to 200
    orc = ob.OracleEnv.from_scenario(sc).expand(nodes, nthreads=8)
    env = gpu_env(sc)
    env.enable_stats(True)
    g = env.expand(nodes, want=WANT)
    assert_expansion_equal(g, orc, exact_cost=True)
    samples, succ = env.last_stats()
    t = ob.OracleEnv.from_scenario(sc).timed(nodes)
    assert samples == t["samples"] and succ == t["successors"]


def test_many_controls_uses_sequential_kernel():
    """|U| = 343 > 256 primitives per node: served by the sequential kernel."""
    from scenarios import control_set

    sc = _custom(3, ACC, control_set(1.5, 7, 3), 48, 0.2, v_max=3.0)
    nodes = sc.frontier(300, seed=9)
    orc = ob.OracleEnv.from_scenario(sc).expand(nodes, nthreads=8)
    g = gpu_env(sc).expand(nodes, want=WANT)
    assert_expansion_equal(g, orc, exact_cost=True)


def _check_packed(env, sc, orc, nodes, drop_inf):
    p = env.expand_packed(nodes, drop_inf=drop_inf)
    nU, D = orc["nU"], sc.Dim
    fields = ["pos", "vel", "acc", "jrk"][: bin(sc.control & 15).count("1")]
    keep_all = (np.arange(nU)[None, :] < orc["count"][:, None])
    if drop_inf:
        keep_all &= ~np.isinf(orc["cost"].reshape(-1, nU))
    np.testing.assert_array_equal(p["count"], keep_all.sum(1))
    assert p["total"] == int(keep_all.sum())
    # every node's records are contiguous at offset[i]; the segments tile [0,total) exactly
    order = np.argsort(p["offset"], kind="stable")
    nz = order[p["count"][order] > 0]
    assert (p["offset"][nz] == np.concatenate([[0], np.cumsum(p["count"][nz])[:-1]])).all()
    sel = np.nonzero(keep_all.reshape(-1))[0]
    # gather the packed records in (node, control) order
    idx = np.concatenate([p["offset"][i] + np.arange(p["count"][i]) for i in range(len(nodes))]) if len(nodes) else []
    exp_state = np.concatenate([orc["succ"][f][sel][:, :D] for f in fields] +
                               ([orc["succ"]["yaw"][sel][:, None]] if sc.control & 16 else []), axis=1)
    assert p["nstate"] == exp_state.shape[1]
    assert p["state"][idx].tobytes() == np.ascontiguousarray(exp_state).tobytes()
    np.testing.assert_array_equal(p["action"][idx], orc["action"][sel].astype(np.uint16))
    np.testing.assert_array_equal(p["key"][idx], orc["key"][sel])
    g, o = p["cost"][idx], orc["cost"][sel]
    np.testing.assert_array_equal(np.isinf(g), np.isinf(o))
    np.testing.assert_allclose(g[~np.isinf(o)], o[~np.isinf(o)], rtol=1e-6, atol=0)
    # the documented reconstruction rule (include/mplx.h, mplx_packed_out): state fields from the
    # record, the first derivative above the state order = 0 + U[action], higher ones 0, yaw 0
    # without a yaw control, t = parent.t + dt  ==>  the full successor Waypoint, bit for bit
    nf = len(fields)
    full = np.zeros(len(idx), dtype=ob.WAYPOINT_DTYPE)
    st = p["state"][idx]
    act = p["action"][idx].astype(np.int64)
    for f_i, name in enumerate(["pos", "vel", "acc", "jrk"]):
        if f_i < nf:
            full[name][:, :D] = st[:, f_i * D:(f_i + 1) * D]
        elif f_i == nf:
            full[name][:, :D] = 0.0 + sc.U[act][:, :D]
    if sc.control & 16:
        full["yaw"] = st[:, nf * D]
    parent = np.repeat(np.arange(len(nodes)), p["count"])
    full["t"] = nodes["t"][parent] + sc.T
    assert full.tobytes() == orc["succ"][sel].tobytes()


def test_packed_stream_matches_oracle():
    """mplx_expand_packed: dense state/cost/action/key records, with and without +inf successors,
    in one chunk and across more pipeline chunks than there are buffer sets (MPLX_PACK_CHUNK_LOG2=18:
    60k nodes x 27 = 6.2 chunks of 2^18 slots over 4 buffer sets)."""
    import os

    import scenarios as S

    for sc, n in ((S.scaled(S.cfg_headline(), 96), 60000), (S.scaled(S.cfg3(), 64), 9000), (S.scaled(S.cfg4(), 64), 3000)):
        nodes = sc.frontier(n, seed=21)
        orc = ob.OracleEnv.from_scenario(sc).expand(nodes, nthreads=8, lattice=False)
        env = gpu_env(sc)
        for chunk_log2 in (None, "18"):
            if chunk_log2 is None:
                os.environ.pop("MPLX_PACK_CHUNK_LOG2", None)
            else:
                os.environ["MPLX_PACK_CHUNK_LOG2"] = chunk_log2
            try:
                for drop in (False, True):
                    _check_packed(env, sc, orc, nodes, drop)
            finally:
                os.environ.pop("MPLX_PACK_CHUNK_LOG2", None)
    # 2-D, pageable buffers, tiny and empty batches
    c = fixtures.corridor()
    from scenarios import Scenario

    sc = Scenario("corridor", tuple(int(x) for x in c["dim"]), c["res"], tuple(c["origin"]), ACC, fixtures.U_2d(),
                  v_max=1.0, a_max=1.0)
    sc._grid = c["grid"]
    env = gpu_env(sc)
    for n in (0, 1, 5, 700):
        nodes = sc.frontier(n, seed=3) if n else np.zeros(0, dtype=ob.WAYPOINT_DTYPE)
        orc = ob.OracleEnv.from_scenario(sc).expand(nodes, lattice=False)
        p = env.expand_packed(nodes, drop_inf=True, pinned=False)
        if n:
            _check_packed(env, sc, orc, nodes, True)
        else:
            assert p["total"] == 0


def test_small_batches_zero_copy_equal_large_batches():
    """mplx_expand runs batches of <= 4096 successor slots zero-copy (the kernel reads and writes pinned host
    memory, csrc/mplx_api.cu); larger ones go through the staged copies.  The same nodes must give the same
    records either way, with pageable and with pinned caller arrays."""
    import scenarios as S

    for sc in (S.scaled(S.cfg_headline(), 64), S.scaled(S.cfg3(), 64)):
        nodes = sc.frontier(3000, seed=3)
        env = gpu_env(sc)
        big = env.expand(nodes, want=WANT)  # 3000 * |U| slots: staged path
        nU = len(sc.U)
        for pinned in (False, True):
            for lo, m in ((0, 1), (17, 4096 // nU), (2900, 7)):
                g = env.expand(nodes[lo:lo + m], want=WANT, pinned=pinned)
                np.testing.assert_array_equal(g.count, big.count[lo:lo + m])
                for i in range(m):
                    k = int(g.count[i])
                    a, b = i * nU, (lo + i) * nU
                    for name in ("succ", "cost", "action", "key"):
                        x, y = getattr(g, name)[a:a + k], getattr(big, name)[b:b + k]
                        assert x.tobytes() == y.tobytes(), (name, pinned, lo, m, i)
