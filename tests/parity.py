"""Comparison helper: CUDA expansion (libmplx) vs the CPU oracle on the same nodes."""
import numpy as np

COST_RTOL = 1e-6  # north_star: edge costs within 1e-6 relative; +inf must match exactly


def assert_expansion_equal(gpu, orc, exact_cost=False):
    """gpu: motion_primitive_library_b200.env.Expansion ; orc: dict from OracleEnv.expand."""
    nU = orc["nU"]
    assert gpu.nU == nU
    np.testing.assert_array_equal(gpu.count, orc["count"])
    n = gpu.count.size
    idx = (np.arange(n)[:, None] * nU + np.arange(nU)[None, :])
    valid = (np.arange(nU)[None, :] < orc["count"][:, None])
    sel = idx[valid]
    if gpu.action is not None:
        np.testing.assert_array_equal(gpu.action[sel], orc["action"][sel])
    if gpu.succ is not None:
        a = gpu.succ[sel].view(np.uint64).reshape(sel.size, -1)
        b = orc["succ"][sel].view(np.uint64).reshape(sel.size, -1)
        bad = np.nonzero((a != b).any(1))[0]
        assert bad.size == 0, f"{bad.size} successor waypoints differ bitwise; first: gpu={gpu.succ[sel][bad[0]]} orc={orc['succ'][sel][bad[0]]}"
    if gpu.key is not None:
        np.testing.assert_array_equal(gpu.key[sel], orc["key"][sel])
    if gpu.lattice is not None and orc.get("lattice") is not None:
        np.testing.assert_array_equal(gpu.lattice[sel], orc["lattice"][sel])
    if gpu.cost is not None:
        g, o = gpu.cost[sel], orc["cost"][sel]
        np.testing.assert_array_equal(np.isinf(g), np.isinf(o))
        fin = ~np.isinf(o)
        if exact_cost:
            np.testing.assert_array_equal(g[fin], o[fin])
        else:
            np.testing.assert_allclose(g[fin], o[fin], rtol=COST_RTOL, atol=0)
    return dict(nodes=n, successors=int(sel.size), finite=int((~np.isinf(orc["cost"][sel])).sum()))
