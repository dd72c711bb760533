"""Host-side logic on the CPU: the product's planner (mpl_host.hpp: heap, state space, A*,
recoverTraj) driven by the oracle env.  config 1 = test/test_planner_2d.cpp on data/corridor.yaml."""
import numpy as np

import fixtures
import oracle_bindings as ob
import planner_bindings as pb

ACC = 0x03


def corridor_args(**kw):
    c = fixtures.corridor()
    return c, pb.make_args(2, ACC, c["grid"], c["dim"], c["origin"], c["res"], fixtures.U_2d(),
                           start=dict(pos=c["start"]), goal=dict(pos=c["goal"]), v_max=1.0, a_max=1.0, T=1.0, w=10.0,
                           eps=1.0, tol_pos=0.5, **kw)


def test_corridor_plan_is_consistent():
    c, args = corridor_args()
    r = pb.plan_oracle(args)
    assert r["valid"] == 1
    assert r["n_actions"] > 20 and r["expanded"] > 100
    assert r["n_closed"] == r["expanded"]  # A* closes one state per iteration; no re-expansion with a consistent heuristic
    assert len(np.unique(r["closed"])) == r["n_closed"]
    # replay the recovered actions through the oracle: reaches the goal box with the same cost
    env = ob.OracleEnv(2, ACC, fixtures.U_2d(), c["grid"], c["dim"], c["origin"], c["res"], T=1.0, w=10.0, v_max=1.0,
                       a_max=1.0)
    node = ob.wp(c["start"], vel=(0, 0))
    total = 0.0
    for a in r["actions"]:
        s = env.get_succ(node)
        j = list(s["action"]).index(a)
        assert np.isfinite(s["cost"][j])
        total += s["cost"][j]
        node = s["succ"][j]
    assert abs(total - r["cost"]) < 1e-9 * max(1.0, total)
    assert np.abs(node["pos"][:2] - c["goal"]).max() <= 0.5
    # lower bound: w * Linf distance / v_max (the heuristic is admissible)
    assert r["cost"] >= 10.0 * np.abs(c["goal"] - c["start"]).max() / 1.0


def test_max_expand_and_blocked_start():
    c, args = corridor_args(max_num=50)
    r = pb.plan_oracle(args)
    assert r["valid"] == 0 and r["expanded"] == 50 and np.isinf(r["cost"])
    c, args = corridor_args()
    args.start.pos[0], args.start.pos[1] = 0.01, -4.99  # inside the wall
    r = pb.plan_oracle(args)
    assert r["valid"] == 0 and r["expanded"] == 0


def test_priority_queue_order_property():
    """Popped f-values are non-decreasing for a consistent heuristic (eps = 1)."""
    c, args = corridor_args(max_num=400)
    r1 = pb.plan_oracle(args)
    c, args = corridor_args(max_num=401)
    r2 = pb.plan_oracle(args)
    # expanding one more node only ever adds to the closed set
    assert set(r1["closed"].tolist()) <= set(r2["closed"].tolist()) and r2["n_closed"] == r1["n_closed"] + 1
