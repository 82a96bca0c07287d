"""The two BASELINE.json workloads that are not the headline metric, as library calls bench.py and tools/ share:
configs[4] (4 M rays against the 1 M-entity tree) and configs[3] (8192 independent arena.sav worlds, sharded).
Host-side plumbing: every number comes from CUDA events on the world's own stream."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


# ---------------------------------------------------------------- configs[4]: batched raycast

def _splitmix32_first(seed):
    """first output of splitmix32 (random.cpp:18-24) for an array of seeds"""
    z = (seed.astype(np.uint64) + np.uint64(0x9e3779b9)) & np.uint64(0xffffffff)
    z ^= z >> np.uint64(16); z = (z * np.uint64(0x21f0aaad)) & np.uint64(0xffffffff)
    z ^= z >> np.uint64(15); z = (z * np.uint64(0x735a2d97)) & np.uint64(0xffffffff)
    z ^= z >> np.uint64(15)
    return z


def make_rays(n_rays, n_ents):
    """ray i starts at entity (i * 2654435761) mod N, look = 3 draws in [-64, 64]"""
    i = np.arange(n_rays, dtype=np.uint64)
    origin = ((i * np.uint64(2654435761)) & np.uint64(0xffffffff)) % np.uint64(n_ents)
    look = np.zeros((n_rays, 3), dtype=np.int32)
    for k in range(3):
        s = (np.uint64(4242) + i + np.uint64(k) * np.uint64(0x9e3779b9)) & np.uint64(0xffffffff)
        look[:, k] = (_splitmix32_first(s) % np.uint64(129)).astype(np.int32) - 64
    look[np.abs(look).sum(1) == 0] = (0, 0, 64)
    return origin.astype(np.int32), look


def raycast_sweep(w, n_rays, reps=3):
    """rays and results resident in HBM; returns (best device ms, winners on the host, origin, look, tree build ms)"""
    import torch
    live = w.count()[0]
    origin, look = make_rays(n_rays, live)
    dev = torch.device("cuda", torch.cuda.current_device())
    d_o = torch.from_numpy(origin).to(dev)
    d_l = torch.from_numpy(look).to(dev)
    d_w = torch.empty(n_rays, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(w.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        w.raycast_build()
        e1.record(stream)
    e1.synchronize()
    build_ms = e0.elapsed_time(e1)
    times = []
    for _ in range(reps + 1):
        with torch.cuda.stream(stream):
            e0.record(stream)
            w.raycast_device(d_o.data_ptr(), d_l.data_ptr(), n_rays, d_w.data_ptr())
            e1.record(stream)
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    return min(times[1:]), d_w.cpu().numpy(), origin, look, build_ms


# ---------------------------------------------------------------- configs[3]: batched arena worlds

def arena_kicks(n_ents, world, max_v):
    rng = np.random.RandomState((world * 7919 + 12345) % (2 ** 31))
    return rng.randint(-max_v, max_v + 1, size=(n_ents, 3)).astype(np.int32)


def arena_batch(level_path, world_ids, kick_max=150):
    """one batched bt_world_desc holding a copy of the level per world id, each with its own seeded velocity kick
    (the device has no player table: without players a shipped level is at rest)"""
    from . import concat_worlds
    from .savfile import load_sav
    base, _ = load_sav(level_path)
    # every handler of the level stays, tick handlers included (respawners, guns nobody holds: they have device bodies, so
    # bt_world_tick runs the whole doWholeStep of a world without players) -- except the scoreboard's: without players
    # it calls every round a tie after 70 ticks, erases the world and asks the HOST to reload the level
    h = base.a["handlers"].reshape(-1, 10)
    h[:, 1][h[:, 1] == 27] = 0   # TICK_SCOREBOARD
    n1 = base.n
    batch = concat_worlds([base] * len(world_ids))
    for k, g in enumerate(world_ids):
        batch.a["d_vel"][3 * n1 * k: 3 * n1 * (k + 1)] = arena_kicks(n1, g, kick_max).reshape(-1)
    return batch, n1


def arena_world(batch, device):
    from . import World
    w = World(None, device=device, max_ents=batch.n, max_sliders=int(batch.desc.n_sliders),
              max_wires=max(1, int(batch.desc.n_wires)), max_worlds=int(batch.desc.n_worlds))
    w.upload(batch)  # strict: every collision-path handler of the level has a device body
    return w
