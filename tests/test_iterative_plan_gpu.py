"""MapPlanner::iterativePlan end to end on the GPU env (tunnels built by mplx_set_search_region_path,
expansion by libmplx) against the same host planner with the CPU oracle env, which
tests/test_iterative_plan_vs_ref.py pins to the reference's own iterativePlan.

Passing on the device since round 1's round-end run (observed again in round 2)."""
import numpy as np
import pytest

import fixtures
import planner_bindings as pb
from test_iterative_plan_vs_ref import same

pytestmark = pytest.mark.gpu
ACC = 0x03


@pytest.mark.parametrize("radius,speculate", [((0.5, 0.5), 1), ((0.15, 0.15), 8)])
def test_corridor_tunnel_replanning(radius, speculate):
    c = fixtures.corridor()
    a = pb.make_args(2, ACC, c["grid"], c["dim"], c["origin"], c["res"], fixtures.U_2d(), start=dict(pos=c["start"]),
                     goal=dict(pos=c["goal"]), v_max=1.0, a_max=1.0, speculate=speculate)
    orc = pb.iterative_oracle(a, radius, 3)
    assert orc[0]["valid"] == 1 and orc[1]["ok"] == 1
    got = pb.iterative_plan(a, radius, 3)
    same(got, orc)
    assert got[1]["iterations"] == orc[1]["iterations"]


def test_voxel_map_tunnel():
    import scenarios as S

    sc = S.scaled(S.cfg_headline(), 64)
    nodes = sc.frontier(16, seed=4, max_steps=0)
    a = pb.make_args(3, sc.control, sc.grid(), sc.dim_cells, sc.origin, sc.res, sc.U, start=dict(pos=nodes["pos"][0]),
                     goal=dict(pos=nodes["pos"][1]), v_max=sc.v_max, a_max=sc.a_max, max_num=4000, speculate=16)
    orc = pb.iterative_oracle(a, (0.6, 0.6, 0.4), 3)
    assert orc[0]["valid"] == 1
    same(pb.iterative_plan(a, (0.6, 0.6, 0.4), 3), orc)
