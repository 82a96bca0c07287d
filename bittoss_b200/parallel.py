"""Multi-GPU layer: a single world does not shard (its resolution loop is a global
fixed-point iteration over chains of pending collisions, ent.cpp:631-672, 866-905), so N
GPUs run N disjoint sets of independent worlds, RL-environment style.  There is no
collective on the tick path; the only exchange is the final gather of per-world state
hashes.  Works with any torch.distributed backend (nccl on the GPUs, gloo in CPU tests)."""
from typing import List, Sequence, Tuple


def shard_range(n_worlds: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of world ids [begin, end) owned by `rank`."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(n_worlds, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_hashes(local_hashes: Sequence[int], n_worlds: int, device=None) -> List[int]:
    """All ranks end up with the full list of per-world 64-bit state hashes, ordered by world id.
    Each rank passes the hashes of its shard_range()."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(h) for h in local_hashes]
    ws, rank = dist.get_world_size(), dist.get_rank()
    longest = max(shard_range(n_worlds, r, ws)[1] - shard_range(n_worlds, r, ws)[0] for r in range(ws))
    # int64 carries the 64-bit pattern (two's complement); pad to the longest shard
    vals = [(h - (1 << 64)) if h >= (1 << 63) else h for h in local_hashes]
    buf = torch.zeros(longest, dtype=torch.int64, device=device)
    if vals:
        buf[: len(vals)] = torch.tensor(vals, dtype=torch.int64, device=device)
    out = [torch.zeros_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf)
    full = []
    for r in range(ws):
        b, e = shard_range(n_worlds, r, ws)
        full.extend(int(x) & ((1 << 64) - 1) for x in out[r][: e - b].tolist())
    return full


def reduce_time_and_work(ms: float, units: float, device=None) -> Tuple[float, float]:
    """max over ranks of the device time, sum over ranks of the work: whole-job throughput."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return ms, units
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    u = torch.tensor([units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
