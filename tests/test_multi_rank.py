"""N>1 path on CPU: world_size-2 gloo.  Each rank steps its shard of a batch of independent
worlds (with the oracle standing in for the GPU engine, which this container cannot run) and
the final hash gather must reproduce the single-process result."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

NAMES = ["arena_kick", "diner_kick", "stack512", "arena_kick", "tower_kick"]
TICKS = 25


def _worker(rank, world_size, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from util import load_scenario
    from oracle.pybind import OracleLib
    from bittoss_b200.parallel import gather_hashes, reduce_time_and_work, shard_range
    import bittoss_b200
    b, e = shard_range(len(NAMES), rank, world_size)
    mine = [load_scenario(n)[0] for n in NAMES[b:e]]
    local = []
    if mine:
        ow = OracleLib().create(bittoss_b200.concat_worlds(mine))
        for _ in range(TICKS):
            ow.step(31)
        local = ow.hash()
    full = gather_hashes(local, len(NAMES))
    t, u = reduce_time_and_work(10.0 + rank, float(len(mine)))
    if rank == 0:
        q.put((full, t, u))
    dist.destroy_process_group()


def test_shard_range_partitions_everything():
    from bittoss_b200.parallel import shard_range
    for n in (0, 1, 5, 8192):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_hash_gather_matches_single_process():
    from util import load_scenario
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, t, u = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [int(load_scenario(n)[1][TICKS - 1]) for n in NAMES]  # the reference's golden hashes
    assert full == want
    assert t == 11.0 and u == float(len(NAMES))
