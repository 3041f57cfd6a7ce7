"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: shard ranges tile the path range, keep
antithetic pairs together, and the sum-all-reduce of per-rank result vectors reproduces the single-rank
result.  The per-rank 'engine' here is the numpy oracle (the GPU engine needs a device)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from compfinance_b200.dist import allreduce_sum_, shard_range


@pytest.mark.parametrize("n,world,anti", [(1 << 20, 8, False), (1 << 20, 3, True), (1000, 2, True), (5, 4, False),
                                          (1 << 22, 8, True), (777, 2, False)])
def test_shards_tile_the_range(n, world, anti):
    nxt = 0
    for r in range(world):
        first, count = shard_range(n, r, world, antithetic=anti)
        assert first == nxt and count >= 0
        if anti and r < world - 1:
            assert first % 2 == 0 and count % 2 == 0
        nxt = first + count
    assert nxt == n


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import restate as R
    spots = np.arange(80.0, 121.0, 10.0)
    times = np.array([0.5, 1.0])
    vols = 0.2 + 0.0 * spots[:, None] + 0.01 * times[None, :]
    ptl = R.uoc_timeline(0.25, 1.0 / 52)
    tab = R.DupireTables(100.0, spots, times, vols, 0.25, ptl)
    first, count = shard_range(n, rank, world, antithetic=True, align=64)
    o = R.dupire_uoc_run(tab, dict(strike=100.0, barrier=115.0, smooth=0.01), ("mrg32k3a", 12345, 12346), first, count, [1.0, 0.5])
    vec = torch.from_numpy(np.concatenate([o["payoffs"].sum(0), [o["agg"].sum(), o["spot_adj"]], o["ybar"].ravel()]))
    allreduce_sum_(vec)
    if rank == 0:
        q.put(vec.numpy())
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_run():
    n, world = 1024, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from oracle import restate as R
    spots = np.arange(80.0, 121.0, 10.0)
    times = np.array([0.5, 1.0])
    vols = 0.2 + 0.0 * spots[:, None] + 0.01 * times[None, :]
    tab = R.DupireTables(100.0, spots, times, vols, 0.25, R.uoc_timeline(0.25, 1.0 / 52))
    o = R.dupire_uoc_run(tab, dict(strike=100.0, barrier=115.0, smooth=0.01), ("mrg32k3a", 12345, 12346), 0, n, [1.0, 0.5])
    want = np.concatenate([o["payoffs"].sum(0), [o["agg"].sum(), o["spot_adj"]], o["ybar"].ravel()])
    assert np.max(np.abs(got - want)) < 1e-9 * max(1.0, np.max(np.abs(want)))
