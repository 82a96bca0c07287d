"""BASELINE.json configs[3] shape: thousands of independent arena.sav worlds in ONE device
world per GPU (world id in the Morton key and node ranges), physics on the device.

Without the host-side tick handlers an arena world is at rest, so every world gets its own
seeded velocity kick (uVel equivalent, carried as d_vel in the upload) to set it in motion.
Reports world-ticks/s and entity-ticks/s (CUDA events on the world's stream) and, per world,
that the batched run equals the same world run alone (hash equality on a sample).

    python tools/batched_arena_bench.py --worlds 8192 --ticks 200 [--level arena.sav]
Under torchrun each rank takes parallel.shard_range(worlds) and the hashes are gathered."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def kicks(n_ents, world, tick, max_v):
    rng = np.random.RandomState((world * 7919 + tick * 104729 + 12345) % (2 ** 31))
    return rng.randint(-max_v, max_v + 1, size=(n_ents, 3)).astype(np.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worlds", type=int, default=8192)
    ap.add_argument("--ticks", type=int, default=200)
    ap.add_argument("--kick-max", type=int, default=150)
    ap.add_argument("--level", default=os.path.join(ROOT, "oracle", "_ref", "data", "arena.sav"))
    ap.add_argument("--check", type=int, default=4)
    ap.add_argument("--patch-shift", type=int, default=-1)
    ap.add_argument("--small", type=int, default=1, help="0: force the LBVH path instead of the small-world pair search")
    ap.add_argument("--horizon", type=int, default=-1)
    ap.add_argument("--patch", type=int, default=1)
    ap.add_argument("--margin", type=int, default=-1)
    ap.add_argument("--profile", action="store_true", help="extra pass with CUDA events around every launch")
    args = ap.parse_args()
    import torch
    import bittoss_b200
    from bittoss_b200.parallel import gather_hashes, reduce_time_and_work, shard_range
    from bittoss_b200.savfile import load_sav

    rank, ws = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if ws > 1:
        import torch.distributed as dist
        torch.cuda.set_device(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    b, e = shard_range(args.worlds, rank, ws)
    base, info = load_sav(args.level)
    base.a["handlers"][1::10] = 0  # tick / tickHeld ids stay on the host: not run here
    base.a["handlers"][2::10] = 0
    n1 = base.n

    def make(world_ids):
        batch = bittoss_b200.concat_worlds([base] * len(world_ids))
        for k, g in enumerate(world_ids):
            batch.a["d_vel"][3 * n1 * k: 3 * n1 * (k + 1)] = kicks(n1, g, 0, args.kick_max).reshape(-1)
        w = bittoss_b200.World(None, device=lr, max_ents=batch.n, max_sliders=int(batch.desc.n_sliders),
                               max_wires=max(1, int(batch.desc.n_wires)), max_worlds=len(world_ids))
        w.set_option("allow_unsupported", 1)
        if args.horizon >= 0:
            w.set_option("cache_horizon", args.horizon)
        if args.patch_shift >= 0:
            w.set_option("patch_shift", args.patch_shift)
        w.set_option("pair_patch", args.patch)
        w.set_option("small_worlds", args.small)
        if args.margin >= 0:
            w.set_option("cache_margin", args.margin)
        w.upload(batch)
        return w

    def run(w, world_ids, timed):
        total_ms = w.tick_timed(args.ticks) if timed else (w.tick(args.ticks) or 0.0)
        return total_ms, w.hash()

    ids = list(range(b, e))
    w = make(ids)
    ms, hashes = run(w, ids, True)
    st = w.stats()
    w.close()
    if args.profile and rank == 0:
        wp = make(ids)
        wp.profile_enable(True)
        wp.tick(args.ticks)
        for k, (n_l, k_ms) in sorted(wp.profile().items(), key=lambda kv: -kv[1][1]):
            if n_l:
                print("   %-16s %8.4f ms/tick  %6.2f launches/tick" % (k, k_ms / args.ticks, n_l / args.ticks), file=sys.stderr)
        wp.close()
    ms_max, et = reduce_time_and_work(ms, float(st["entity_ticks"]), device="cuda" if ws > 1 else None)
    all_hashes = gather_hashes(hashes, args.worlds, device="cuda" if ws > 1 else None)
    ok = True
    if rank == 0:
        for g in ids[: args.check]:  # the same world alone must give the same hash
            w1 = make([g])
            _, h1 = run(w1, [g], False)
            w1.close()
            ok = ok and h1[0] == all_hashes[g]
        print(json.dumps({
            "metric": "world_ticks_per_sec", "value": args.worlds * args.ticks / (ms_max * 1e-3), "unit": "world-ticks/s",
            "entity_ticks_per_sec": et / (ms_max * 1e-3), "n_gpus": ws, "worlds": args.worlds, "ticks": args.ticks,
            "entities_per_world": n1, "ms_per_tick": ms_max / args.ticks, "level": os.path.basename(args.level),
            "iterations_per_tick": st["iterations"] / max(1, st["ticks"]), "crushes": st["crushes"],
            "unsupported_events": st["unsupported_events"], "pair_builds": st["pair_builds"], "pair_patches": st["pair_patches"], "tree_builds": st["tree_builds"], "distinct_hashes": len(set(all_hashes)),
            "alone_equals_batched": ok, "checked_worlds": args.check}))
    if dist is not None:
        dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
