"""Synthetic-input generators (host logic, CPU)."""
import numpy as np

import scenarios as S


def test_potential_field_fast_path_equals_literal_stencil():
    """potential_from_map (per-layer distance transform) == the literal createMask/updatePotentialMap
    stamping rule of src/mpl_planner/map_planner.cpp:286-391."""
    rng = np.random.default_rng(0)
    for dims, res, rad in (((40, 36, 30), 0.1, (0.5, 0.5, 0.3)), ((32, 32, 32), 0.25, (1.0, 1.0, 0.5)),
                           ((30, 30, 30), 0.1, (0.35, 0.35, 0.45))):
        g = (rng.random(int(np.prod(dims))) < 0.01).astype(np.int8) * 100
        np.testing.assert_array_equal(S.potential_from_map(g, dims, res, rad), S.potential_from_map_stencil(g, dims, res, rad))


def test_control_set_order_matches_reference_loops():
    """test/test_planner_2d.cpp:52-53: dx outer, dy inner; yaw innermost (test_planner_2d_with_yaw.cpp:52-57)."""
    U = S.control_set(0.5, 3, 2)
    assert U.tolist() == [[-0.5, -0.5], [-0.5, 0], [-0.5, 0.5], [0, -0.5], [0, 0], [0, 0.5], [0.5, -0.5], [0.5, 0], [0.5, 0.5]]
    Uy = S.control_set(1.0, 3, 3, yaw_rates=(-0.5, 0.0, 0.5))
    assert Uy.shape == (81, 4) and Uy[0].tolist() == [-1, -1, -1, -0.5] and Uy[1].tolist() == [-1, -1, -1, 0.0]
    assert not np.signbit(U).any() or (U[np.signbit(U)] != 0).all()  # no negative zeros


def test_frontier_nodes_are_free_reachable_lattice_states():
    sc = S.scaled(S.cfg_headline(), 64)
    f = sc.frontier(500, seed=1)
    g = sc.grid().reshape(64, 64, 64)
    cell = np.floor((f["pos"] - np.asarray(sc.origin)) / sc.res).astype(int)
    assert ((cell >= 0) & (cell < 64)).all()
    assert (g[cell[:, 2], cell[:, 1], cell[:, 0]] == 0).all()
    assert np.abs(f["vel"]).max() <= sc.v_max
    np.testing.assert_array_equal(sc.frontier(500, seed=1), f)  # deterministic
