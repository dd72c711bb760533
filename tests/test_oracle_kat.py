"""Known-answer tests that pin the CPU oracle (analytic values).

The reference ships no golden vectors for this path (its tests are crash-only,
SURVEY.md §4), so these analytic KATs + the oracle/_ref cross-check (test_oracle_vs_ref.py)
are what the oracle is anchored on.
"""
import math

import numpy as np
import pytest

import fixtures
import oracle_bindings as ob

ACC, JRK, VEL, SNP, ACCxYAW = 0x03, 0x07, 0x01, 0x0F, 0x13


@pytest.fixture(scope="module")
def corridor_env():
    c = fixtures.corridor()
    return c, ob.OracleEnv(2, ACC, fixtures.U_2d(), c["grid"], c["dim"], c["origin"], c["res"], T=1.0, w=10.0,
                           v_max=1.0, a_max=1.0)


def test_corridor_fixture_matches_survey():
    c = fixtures.corridor()
    assert tuple(c["dim"]) == (799, 199) and c["res"] == 0.05
    assert (c["raw"] > 0).sum() == 35067 and (c["raw"] < 0).sum() == 123934
    # start (2.5,-3.5) -> cell (50,30), free  
    pn = [int(round((c["start"][k] - c["origin"][k]) / c["res"] - 0.5)) for k in range(2)]
    # python round is half-even; (2.5-0)/0.05-0.5 = 49.5 -> C round gives 50
    pn = [int(math.floor(((c["start"][k] - c["origin"][k]) / c["res"] - 0.5) + 0.5)) for k in range(2)]
    assert pn == [50, 30]
    assert c["grid"][pn[0] + 799 * pn[1]] == 0


def test_first_expansion_of_test_planner_2d(corridor_env):
    """start (2.5,-3.5) at rest, ACC, U={-.5,0,.5}^2, T=1, w=10, v_max=1."""
    c, env = corridor_env
    r = env.get_succ(ob.wp(c["start"], vel=(0, 0)))
    assert list(r["action"]) == [0, 1, 2, 3, 5, 6, 7, 8]  # control (0,0) dropped: self loop (Q3)
    exp_pos = [(2.25, -3.75), (2.25, -3.5), (2.25, -3.25), (2.5, -3.75), (2.5, -3.25), (2.75, -3.75), (2.75, -3.5),
               (2.75, -3.25)]
    exp_vel = [(-0.5, -0.5), (-0.5, 0), (-0.5, 0.5), (0, -0.5), (0, 0.5), (0.5, -0.5), (0.5, 0), (0.5, 0.5)]
    exp_cost = [10.5, 10.25, 10.5, 10.25, 10.25, 10.5, 10.25, 10.5]
    exp_lat = [(225, -5, -375, -5), (225, -5, -350, 0), (225, -5, -325, 5), (250, 0, -375, -5), (250, 0, -325, 5),
               (275, 5, -375, -5), (275, 5, -350, 0), (275, 5, -325, 5)]
    for s in range(8):
        assert tuple(r["succ"]["pos"][s][:2]) == exp_pos[s]
        assert tuple(r["succ"]["vel"][s][:2]) == exp_vel[s]
        assert r["cost"][s] == exp_cost[s]
        assert tuple(r["lattice"][s][:4]) == exp_lat[s]
        assert r["succ"]["t"][s] == 1.0
        # acc of tn is the evaluated derivative = u even though acc is not part of the ACC state
        assert tuple(r["succ"]["acc"][s][:2]) == tuple(fixtures.U_2d()[r["action"][s]])
    assert len(set(r["key"].tolist())) == 8
    # each moving primitive: max_v=0.5 -> n=10 -> 11 samples (Q8)
    assert ob.lib().orc_last_samples() == 8 * 11


def test_sample_count_table():
    """for(t=0;t<T;t+=T/n): n+1 iterations for 67 of the 124 n in [5,128]."""
    L = ob.lib()
    plus1 = [n for n in range(5, 129) if L.orc_sample_count(1.0, n) == n + 1]
    exact = [n for n in range(5, 129) if L.orc_sample_count(1.0, n) == n]
    assert len(plus1) + len(exact) == 124
    assert len(plus1) == 67
    for n in (6, 7, 10, 13, 14, 15, 19, 22, 23, 24, 26):
        assert n in plus1
    for n in (5, 8, 9, 11, 12, 16, 20):
        assert n in exact
    # python restatement of the same loop agrees for every n
    for n in range(5, 129):
        k, t, dt = 0, 0.0, 1.0 / n
        while t < 1.0:
            k += 1
            t += dt
        assert k == L.orc_sample_count(1.0, n)


def test_lattice_rounding_ties():
    """std::round(x/res) ties on real lattices."""
    grid = np.zeros(16, dtype=np.int8)
    env = ob.OracleEnv(2, ACC, fixtures.U_2d(), grid, (4, 4), (0, 0), 1.0)
    import ctypes as C

    def lat(pos, vel):
        w = ob.wp(pos, vel=vel)
        arr = np.zeros(1, dtype=ob.WAYPOINT_DTYPE)
        arr[0] = w
        out = np.zeros(13, dtype=np.int32)
        n = C.c_int32()
        ob.lib().orc_hash(C.byref(env.e), arr.ctypes.data, out.ctypes.data, C.byref(n))
        assert n.value == 4
        return out[:4].tolist()

    assert lat((0.125, 0.0), (0.25, 0.75)) == [13, 3, 0, 8]  # 12.5->13, 2.5->3, 7.5->8
    assert lat((0.0, 0.0), (0.35, 0.15)) == [0, 3, 0, 1]  # 3.4999999999999996->3, 1.4999999999999998->1
    assert lat((-0.125, 0.0), (-0.25, -0.75)) == [-13, -3, 0, -8]  # half away from zero


def test_jrk_velocity_extremum():
    """JRK: interior vel root at -c3/c2 = -a/u (primitive.h:152-162 via math.h:126-128)."""
    grid = np.zeros(8, dtype=np.int8)
    U = np.array([[-2.0, 0.0, 0.0]])
    env = ob.OracleEnv(3, JRK, U, grid, (2, 2, 2), (0, 0, 0), 1.0)
    w = ob.wp((0, 0, 0), vel=(0.5, 0, 0), acc=(1.0, 0, 0))
    arr = np.zeros(1, dtype=ob.WAYPOINT_DTYPE)
    arr[0] = w
    import ctypes as C

    # v(t) = -2/2 t^2 + 1 t + 0.5 ; root t=0.5 ; v(0.5) = -0.25+0.5+0.5 = 0.75 ; v(0)=0.5 ; v(1)=0.5
    assert ob.lib().orc_max_vel(C.byref(env.e), arr.ctypes.data, 0, 0) == 0.75
    assert ob.lib().orc_max_vel(C.byref(env.e), arr.ctypes.data, 0, 1) == 0.0


def test_acc_primitive_from_rest_cost():
    """ACC from rest, u=(0.5,0), T=1: end pos +0.25, vel 0.5, J=0.25, cost 0.25+10."""
    grid = np.zeros(100 * 100, dtype=np.int8)
    env = ob.OracleEnv(2, ACC, np.array([[0.5, 0.0]]), grid, (100, 100), (0, 0), 0.1, w=10.0)
    r = env.get_succ(ob.wp((5.0, 5.0), vel=(0, 0)))
    assert r["count"][0] == 1
    assert tuple(r["succ"]["pos"][0][:2]) == (5.25, 5.0)
    assert tuple(r["succ"]["vel"][0][:2]) == (0.5, 0.0)
    assert r["cost"][0] == 10.25


def test_collision_outside_and_inf_entries_are_emitted():
    """Q6: colliding / out-of-map but dynamically valid successors come back with +inf."""
    grid = np.zeros(20 * 20, dtype=np.int8)
    grid[10 + 20 * 12] = 100  # cell (10,12)
    env = ob.OracleEnv(2, VEL, np.array([[0.0, 1.0], [0.0, -1.0], [5.0, 0.0]]), grid, (20, 20), (0, 0), 0.1)
    r = env.get_succ(ob.wp((1.05, 0.55)))
    assert r["count"][0] == 3
    assert np.isinf(r["cost"][0])  # runs into (10,12)
    assert np.isinf(r["cost"][1])  # leaves the map through y<0
    assert np.isinf(r["cost"][2])  # leaves through x>=2.0
    # value 99 is NOT occupied (map_util.h:48: == val_occ only)
    grid[10 + 20 * 12] = 99
    env = ob.OracleEnv(2, VEL, np.array([[0.0, 1.0]]), grid, (20, 20), (0, 0), 0.1, w=10.0)
    r = env.get_succ(ob.wp((1.05, 0.55)))
    assert r["cost"][0] == 1.0 + 10.0  # J = u^2 T = 1, + w T


def test_zero_velocity_primitive_skips_collision_check():
    """Q4: curr.pos == tn.pos exactly -> cost = 0 + intrinsic, no traverse even inside an obstacle."""
    grid = np.full(10 * 10, 100, dtype=np.int8)
    U = np.array([[1.0, 0.0]])
    env = ob.OracleEnv(2, ACC, U, grid, (10, 10), (0, 0), 0.1, w=10.0)
    # v=-0.5, u=1: p(1) = 0.5 - 0.5 + p = p
    r = env.get_succ(ob.wp((0.5, 0.5), vel=(-0.5, 0)))
    assert r["count"][0] == 1 and r["cost"][0] == 1.0 + 10.0
    assert ob.lib().orc_last_samples() == 0


def test_potential_and_region_branches():
    grid = np.zeros(40 * 40, dtype=np.int8)
    pot = np.zeros(40 * 40, dtype=np.int8)
    pot[:] = 50
    U = np.array([[1.0, 0.0]])
    env = ob.OracleEnv(2, VEL, U, grid, (40, 40), (0, 0), 0.1, w=10.0, potential=pot, potential_weight=0.5)
    r = env.get_succ(ob.wp((1.05, 1.05)))
    n = ob.lib().orc_last_samples()
    assert n == 11  # max_v=1, res .1 -> n=10 -> 11 samples
    assert r["cost"][0] == pytest.approx(11 * 0.1 * (0.5 * 50) + 1.0 + 10.0, rel=1e-12)
    pot[:] = 100
    assert np.isinf(env.get_succ(ob.wp((1.05, 1.05)))["cost"][0])
    pot[:] = -1  # unknown/negative potential is free, no cost
    assert env.get_succ(ob.wp((1.05, 1.05)))["cost"][0] == 11.0
    region = np.zeros(40 * 40, dtype=np.uint8)
    env2 = ob.OracleEnv(2, VEL, U, grid, (40, 40), (0, 0), 0.1, region=region)
    assert np.isinf(env2.get_succ(ob.wp((1.05, 1.05)))["cost"][0])
    region[:] = 1
    env3 = ob.OracleEnv(2, VEL, U, grid, (40, 40), (0, 0), 0.1, region=region)
    assert env3.get_succ(ob.wp((1.05, 1.05)))["cost"][0] == 11.0


def test_yaw_fov_and_alignment_cost():
    """validate_yaw (primitive.h:503-525) and the wyaw term (env_map.h:122-129)."""
    grid = np.zeros(100 * 100, dtype=np.int8)
    U = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [-1.0, 0.0, 0.0]])
    env = ob.OracleEnv(2, ACCxYAW, U, grid, (100, 100), (0, 0), 0.1, w=10.0, wyaw=1.0, yaw_max=0.7)
    r = env.get_succ(ob.wp((5, 5), vel=(0, 0), yaw=0.0))
    # u=(1,0): end vel (1,0), aligned with yaw 0 -> valid, zero alignment cost
    # u=(0,1): end vel (0,1) is 90deg off -> d=0 < cos(.7) -> rejected ; u=(-1,0): d=-1 -> rejected
    assert list(r["action"]) == [0]
    assert r["cost"][0] == 1.0 + 10.0
    env = ob.OracleEnv(2, ACCxYAW, U, grid, (100, 100), (0, 0), 0.1, w=10.0, wyaw=2.0, yaw_max=-1)
    r = env.get_succ(ob.wp((5, 5), vel=(0, 0), yaw=0.0))
    assert list(r["action"]) == [0, 1, 2]
    # u=(0,1): every sample with |v|>1e-5 adds wyaw*(1-0)*dt ; n=max(5,ceil(1/0.1))=10 -> 11 samples, first has v=0
    assert r["cost"][1] == pytest.approx(2.0 * 10 * 0.1 + 11.0, rel=1e-12)
    assert r["cost"][2] == pytest.approx(2.0 * 2.0 * 10 * 0.1 + 11.0, rel=1e-12)


def test_snp_unsorted_root_early_break():
    """Q7 (primitive.h:155-160): SNP vel roots come unsorted from quad(); a first root >= T
    breaks the scan and hides a later interior root."""
    grid = np.zeros(8, dtype=np.int8)
    # v(t) = c1/6 t^3 + c2/2 t^2 + c3 t + c4 with c1=u. dv/dt = u/2 t^2 + j t + a.
    # u=-4, j=3, a=-0.5: roots of -2t^2+3t-0.5: quad(b=-2,c=3,d=-.5): p=9-4=5;
    #   r0=(-3-sqrt5)/(-4)=1.309 (>=T -> break) ; r1=(-3+sqrt5)/(-4)=0.19098 (interior, never visited)
    U = np.array([[-4.0, 0.0, 0.0]])
    env = ob.OracleEnv(3, SNP, U, grid, (2, 2, 2), (0, 0, 0), 1.0)
    arr = np.zeros(1, dtype=ob.WAYPOINT_DTYPE)
    arr[0] = ob.wp((0, 0, 0), vel=(0.0, 0, 0), acc=(-0.5, 0, 0), jrk=(3.0, 0, 0))
    import ctypes as C

    mv = ob.lib().orc_max_vel(C.byref(env.e), arr.ctypes.data, 0, 0)
    v1 = -4.0 / 6 + 3.0 / 2 - 0.5
    assert mv == pytest.approx(abs(v1), rel=1e-15)  # only the end points were considered
    t = (-3 + math.sqrt(5)) / (-4)
    v_int = abs(-4.0 / 6 * t ** 3 + 1.5 * t * t - 0.5 * t)
    assert v_int < abs(v1) or True  # documented quirk; value check above is the KAT
