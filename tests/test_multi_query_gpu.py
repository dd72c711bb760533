"""config 5: lock-step batched A* (MPL::MultiQueryPlanner, one device launch per iteration for
all live queries) must give every query exactly the result of planning it alone."""
import numpy as np
import pytest

import fixtures
import planner_bindings as pb

pytestmark = pytest.mark.gpu
ACC, JRK = 0x03, 0x07


def test_batch_equals_individual_plans_corridor():
    c = fixtures.corridor()
    rng = np.random.default_rng(1)
    free = np.nonzero(c["grid"].reshape(199, 799) == 0)
    nq = 24
    pick = rng.choice(len(free[0]), size=2 * nq, replace=False)
    pts = np.stack([(free[1][pick] + 0.5) * c["res"] + c["origin"][0], (free[0][pick] + 0.5) * c["res"] + c["origin"][1]], 1)
    starts = np.zeros(nq, dtype=pb.plan_batch.__globals__["WAYPOINT_DTYPE"])
    goals = starts.copy()
    starts["pos"][:, :2], goals["pos"][:, :2] = pts[:nq], pts[nq:]
    base = dict(v_max=1.0, a_max=1.0, max_num=1500)
    args = pb.make_args(2, ACC, c["grid"], c["dim"], c["origin"], c["res"], fixtures.U_2d(), start=dict(pos=pts[0]),
                        goal=dict(pos=pts[1]), **base)
    res, tot = pb.plan_batch(args, starts, goals)
    assert tot["nodes"] == int(res["expanded"].sum()) and tot["iterations"] == int(res["expanded"].max())
    for q in range(nq):
        a = pb.make_args(2, ACC, c["grid"], c["dim"], c["origin"], c["res"], fixtures.U_2d(), start=dict(pos=pts[q]),
                         goal=dict(pos=pts[nq + q]), **base)
        one = pb.plan_gpu(a)
        ref = pb.plan_oracle(a)
        assert res["valid"][q] == one["valid"] == ref["valid"]
        assert res["expanded"][q] == one["expanded"] == ref["expanded"]
        assert res["n_closed"][q] == one["n_closed"] and res["n_actions"][q] == one["n_actions"]
        if one["valid"]:
            assert res["cost"][q] == one["cost"]
            assert res["cost"][q] == pytest.approx(ref["cost"], rel=1e-6)


def test_batch_3d_jrk_voxel_map():
    import scenarios as S

    sc = S.scaled(S.cfg3(), 64)
    nodes = sc.frontier(32, seed=5, max_steps=0)
    starts, goals = nodes[:16].copy(), nodes[16:].copy()
    args = pb.make_args(3, JRK, sc.grid(), sc.dim_cells, sc.origin, sc.res, sc.U, start=dict(pos=starts["pos"][0]),
                        goal=dict(pos=goals["pos"][0]), v_max=sc.v_max, a_max=sc.a_max, max_num=200)
    res, tot = pb.plan_batch(args, starts, goals)
    for q in range(16):
        a = pb.make_args(3, JRK, sc.grid(), sc.dim_cells, sc.origin, sc.res, sc.U, start=dict(pos=starts["pos"][q]),
                         goal=dict(pos=goals["pos"][q]), v_max=sc.v_max, a_max=sc.a_max, max_num=200)
        one = pb.plan_gpu(a)
        assert (res["valid"][q], res["expanded"][q], res["n_closed"][q]) == (one["valid"], one["expanded"], one["n_closed"])
        if one["valid"]:
            assert res["cost"][q] == one["cost"]
