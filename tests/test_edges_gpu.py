"""GPU parity of the stored-edge queries (mplx_edges_is_free / mplx_edges_cells, through the C ABI)
against the CPU oracle, which tests/test_edges_oracle_vs_ref.py pins to the reference's
env_map::is_free(pr) and tests/test_lpastar_vs_ref.py to MapPlanner::getLinkedNodes.
Bar: free flags, cell lists and offsets bit-exact; intrinsic costs bit-exact (no sampling sums)."""
import numpy as np
import pytest

import oracle_bindings as ob
from test_edges_oracle_vs_ref import edges_of
from test_expand_parity_gpu import gpu_env

pytestmark = pytest.mark.gpu

VEL, ACC, JRK, SNP, ACCxYAW = 0x01, 0x03, 0x07, 0x0F, 0x13


def check(sc, nodes, region=None, extra=500, seed=0):
    rng = np.random.default_rng(seed)
    orc = ob.OracleEnv.from_scenario(sc, region=region)
    parents, actions, _ = edges_of(orc, nodes, rng, extra=extra)
    env = gpu_env(sc, region)
    fo, co = orc.edges_is_free(parents, actions)
    fg, cg = env.is_free_edges(parents, actions)
    np.testing.assert_array_equal(fg, fo)
    assert cg.tobytes() == co.tobytes()
    oo, oc = orc.edges_cells(parents, actions)
    go, gc = env.edge_cells(parents, actions)
    np.testing.assert_array_equal(go, oo)
    np.testing.assert_array_equal(gc, oc)
    # the inverted voxel -> edges table: a STABLE sort of the emitted (getIndex, edge) pairs by voxel
    go2, gc2, tv, te = env.edge_cells(parents, actions, table=True)
    np.testing.assert_array_equal(gc2, oc)
    dims = np.asarray(sc.dim_cells, dtype=np.int64)
    strides = np.concatenate([[1], np.cumprod(dims[:-1])])
    ids = ((oc.astype(np.int64) * strides).sum(1)).astype(np.int32)      # int32 wrap-around as getIndex
    owner = np.repeat(np.arange(parents.size, dtype=np.int32), np.diff(oo))
    order = np.argsort(ids, kind="stable")
    np.testing.assert_array_equal(tv, ids[order])
    np.testing.assert_array_equal(te, owner[order])
    assert 0 < fo.sum() < fo.size and oc.shape[0] > parents.size
    return env, parents, actions


@pytest.mark.parametrize("control,dim", [(VEL, 2), (ACC, 2), (JRK, 2), (SNP, 2), (ACCxYAW, 2), (VEL, 3), (ACC, 3),
                                         (JRK, 3), (SNP, 3)])
def test_all_controls_random_states(control, dim):
    from scenarios import Scenario, control_set
    from test_oracle_vs_ref import random_nodes

    rng = np.random.default_rng(200 + control * 10 + dim)
    yaw = bool(control & 16)
    U = control_set(1.0 if (control & 15) < SNP else 4.0, 3, dim, yaw_rates=(-0.4, 0.0, 0.4) if yaw else None)
    cells, res = 40, 0.2
    sc = Scenario("x", (cells,) * dim, res, tuple(-cells * res / 2 for _ in range(dim)), control, U, n_boxes=8,
                  edge_m=(0.6, 1.8), seed=3, v_max=2.5, a_max=3.0, j_max=6.0, yaw_max=0.9 if yaw else -1.0, wyaw=1.5)
    region = (rng.random(cells ** dim) < 0.97).astype(np.uint8) if dim == 2 else None
    check(sc, random_nodes(rng, 400, dim, cells * res / 2, yaw=yaw), region=region, seed=control)


def test_headline_and_jrk_workload_edges():
    import scenarios as S

    for sc, n in ((S.scaled(S.cfg_headline(), 128), 3000), (S.scaled(S.cfg3(), 96), 500)):
        check(sc, sc.frontier(n, seed=9), extra=2000)


def test_capacity_retry_empty_and_errors():
    from motion_primitive_library_b200 import abi
    import scenarios as S
    import ctypes as C

    sc = S.scaled(S.cfg_headline(), 64)
    env, parents, actions = check(sc, sc.frontier(300, seed=2), extra=0)
    # empty batch
    f, c = env.is_free_edges(parents[:0], actions[:0])
    assert f.size == 0 and c.size == 0
    off, cells = env.edge_cells(parents[:0], actions[:0])
    assert off.tolist() == [0] and cells.shape == (0, 3)
    # too small a capacity: fails with the needed total reported
    n = parents.size
    off = np.zeros(n + 1, dtype=np.int64)
    total = C.c_int64(0)
    small = np.zeros((4, 3), dtype=np.int32)
    rc = env._lib.mplx_edges_cells(env._h, parents.ctypes.data, actions.ctypes.data, n, off.ctypes.data,
                                   small.ctypes.data, 4, C.byref(total), None, None)
    assert rc != 0 and total.value > 4 and off[-1] == total.value
    # an action id outside U
    bad = actions.copy()
    bad[3] = 27
    with pytest.raises(abi.MplxError):
        env.is_free_edges(parents, bad)
