"""N>1 host logic on CPU: two gloo ranks shard a list of planning queries, each runs its slice
(here through the CPU planner harness, so no GPU is needed), and the results/counters are gathered.
The GPU path uses the same run_sharded() with NCCL and MultiQueryPlanner (test_multi_query_gpu.py)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _queries(n):
    import fixtures

    c = fixtures.corridor()
    rng = np.random.default_rng(0)
    free = np.nonzero(c["grid"].reshape(199, 799) == 0)
    pick = rng.choice(len(free[0]), size=2 * n, replace=False)
    pts = np.stack([(free[1][pick] + 0.5) * c["res"] + c["origin"][0], (free[0][pick] + 0.5) * c["res"] + c["origin"][1]], 1)
    q = np.zeros(n, dtype=[("start", "f8", 2), ("goal", "f8", 2)])
    q["start"], q["goal"] = pts[:n], pts[n:]
    return c, q


def _plan_slice(c, qs):
    import fixtures
    import planner_bindings as pb

    res = np.zeros(len(qs), dtype=[("valid", "i4"), ("cost", "f8"), ("expanded", "i4")])
    secs = 0.0
    for i, q in enumerate(qs):
        a = pb.make_args(2, 0x03, c["grid"], c["dim"], c["origin"], c["res"], fixtures.U_2d(), start=dict(pos=q["start"]),
                         goal=dict(pos=q["goal"]), v_max=1.0, a_max=1.0, max_num=300)
        r = pb.plan_oracle(a)
        res[i] = (r["valid"], r["cost"], r["expanded"])
        secs += r["seconds"]
    return res, dict(expansions=int(res["expanded"].sum()), seconds_max=secs, queries=len(qs))


def _worker(rank, world, port, n, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from motion_primitive_library_b200.sharding import run_sharded

    c, q = _queries(n)
    allres, counters = run_sharded(q, lambda mine: _plan_slice(c, mine))
    np.save(Path(outdir) / f"res{rank}.npy", allres)
    np.save(Path(outdir) / f"cnt{rank}.npy", np.array([counters["expansions"], counters["queries"], counters["seconds_max"]]))
    dist.destroy_process_group()


def _bcast_worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from motion_primitive_library_b200.sharding import broadcast_array

    rng = np.random.default_rng(5)
    grid = (rng.random((7, 9, 11)) < 0.1).astype(np.int8) * 100          # only rank 0's copy matters
    pot = rng.integers(-1, 101, 50, dtype=np.int64).astype(np.float64)
    got = broadcast_array(grid if rank == 0 else None, src=0)
    got2 = broadcast_array(pot if rank == 0 else None, src=0)
    empty = broadcast_array(np.zeros((0, 3), dtype=np.int32) if rank == 0 else None, src=0)
    np.save(Path(outdir) / f"grid{rank}.npy", got)
    np.save(Path(outdir) / f"pot{rank}.npy", got2)
    assert empty.shape == (0, 3) and empty.dtype == np.int32
    dist.destroy_process_group()


def test_map_broadcast_two_ranks(tmp_path):
    """The set-up collective: rank 0's map reaches every rank bit for bit (dtype and shape included)."""
    mp.spawn(_bcast_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    g0, g1 = np.load(tmp_path / "grid0.npy"), np.load(tmp_path / "grid1.npy")
    assert g0.dtype == np.int8 and g0.shape == (7, 9, 11) and g0.sum() > 0
    np.testing.assert_array_equal(g0, g1)
    np.testing.assert_array_equal(np.load(tmp_path / "pot0.npy"), np.load(tmp_path / "pot1.npy"))


def test_shard_slice_partitions():
    from motion_primitive_library_b200.sharding import shard_slice

    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            idx = np.concatenate([np.arange(n)[shard_slice(n, r, world)] for r in range(world)])
            np.testing.assert_array_equal(idx, np.arange(n))
            sizes = [shard_slice(n, r, world).stop - shard_slice(n, r, world).start for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_slice(4, 2, 2)


def test_two_rank_gloo_matches_single_process(tmp_path):
    n = 9  # odd: ragged shards (4 + 5)
    c, q = _queries(n)
    single, cs = _plan_slice(c, q)
    mp.spawn(_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(tmp_path / f"res{r}.npy")
        np.testing.assert_array_equal(got["valid"], single["valid"])
        np.testing.assert_array_equal(got["expanded"], single["expanded"])
        np.testing.assert_allclose(got["cost"], single["cost"], rtol=0, atol=0)
        cnt = np.load(tmp_path / f"cnt{r}.npy")
        assert cnt[0] == cs["expansions"] and cnt[1] == n and cnt[2] > 0
