#!/usr/bin/env python
"""bench.py — entity-ticks/sec of the per-tick physics hot path on the synthetic
1M-entity box-stack world (BASELINE.json configs[2]), one independent world per GPU.

    python bench.py --gpus N --steps K --warmup W          # the CUDA path (this repo)
    python bench.py --impl reference --steps K --warmup W  # the reference's own CPU tick

A "step" is one whole simulation tick (flushCtrls, flush1, gravity, doPhysics, finishStep)
of the world.  See DESIGN.md §measurement for every field of the JSON line.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "entity_ticks_per_sec"
UNIT = "entity-ticks/s"
BYTES_PER_ENTITY_TICK = 92          # 56 B read + 36 B written per entity per tick
COLLIDE_BYTES_PER_LEAF = 68         # own geometry 48 B + masks/handler/root/world 20 B


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.p, self.first = index, [], None, 0

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        """rows read so far were taken before the load of interest began (nvidia-smi needs a few
        hundred ms to start: it is launched ahead of the timed region and the early rows dropped)"""
        self.first = len(self.rows)

    def since_mark(self):
        return len(self.rows) - self.first

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


REF_HASH_FILE = os.path.join(os.environ.get("TMPDIR", "/tmp"), "bittoss_b200_reference_arm_hashes_%d.json")
GOLDEN_1M = os.path.join(ROOT, "tests", "golden", "stack1m_ref_hashes.npz")


def cpu_baseline(n_cubes, warm_ticks, ticks, settle=0, keep_hashes=False):
    """The UNMODIFIED reference's CPU tick (oracle/_ref) when it was compiled, else the C port."""
    from oracle import pybind
    cores = 1  # the reference tick is single-threaded and non-reentrant (ent.cpp:17, velbox/box.cpp:17-18)
    hashes = []
    if pybind.have_ref():
        with pybind.stdout_to_stderr():  # the reference chats on stdout ("Safety crush!" ...)
            R = pybind.RefLib()
            R.new_world()
            R.make_boxstack(n_cubes, 4, 777, 64, 60)
            for _ in range(settle + warm_ticks):
                R.step(1)
                if keep_hashes:
                    hashes.append(R.hash())
            if keep_hashes:  # tick by tick, so that every tick's state hash can be kept (the hash walk is outside the clock)
                secs, et = 0.0, 0
                for _ in range(ticks):
                    s1, e1 = R.time_steps(1, 1)
                    secs += s1
                    et += e1
                    hashes.append(R.hash())
            else:
                secs, et = R.time_steps(ticks, 1)
        kind = "reference"
    else:
        import bittoss_b200
        wa = bittoss_b200.synth_boxstack(n_cubes, 4, 777, 64, 60)
        ow = pybind.OracleLib().create(wa)
        for _ in range(settle + warm_ticks):
            ow.step(31)
        secs, et = ow.time_steps(ticks)
        kind = "port"
    out = {"value": et / secs, "unit": UNIT, "cores": cores, "kind": kind, "seconds": round(secs, 3),
           "sample": "%d whole ticks of the %d-cube box-stack world after %d settle + %d warm-up ticks, 1 thread"
                     % (ticks, n_cubes, settle, warm_ticks)}
    if keep_hashes:
        out["_hashes"] = hashes
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the SAME world the CUDA arm times (configs[2] at its named size): about 3 s per tick at 1 M cubes on one core, so
    # settle + warm-up + K timed ticks is a few minutes; the per-tick state hashes are left behind for the CUDA arm of
    # the same box to compare its own stream with (`parity.vs_reference_arm` in its line)
    n_cubes = int(args.entities)
    cb = cpu_baseline(n_cubes, args.warmup, args.steps, args.settle, keep_hashes=True)
    hashes = cb.pop("_hashes", [])
    try:
        with open(REF_HASH_FILE % n_cubes, "w") as f:
            json.dump({"cubes": n_cubes, "seed": 777, "height": 4, "push_every": 64, "push_max": 60, "kind": cb["kind"],
                       "hashes": ["%016x" % h for h in hashes]}, f)
    except OSError:
        pass
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cb["seconds"] / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "synthetic 1M-entity integer AABB box-stack world with gravity and pushes "
                                   "(BASELINE.json configs[2]), one independent world per GPU",
                       "cubes": n_cubes, "height": 4, "push_every": 64, "push_max": 60, "settle_ticks": args.settle,
                       "sample": cb["sample"], "state_hashes_kept": len(hashes)},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def parity_record(bittoss_b200, args, device, checkpoints):
    """Bit-exactness of THIS run against the unmodified reference at the named size (rank 0: seed 777).
    `checkpoints` = {ticks done: state hash} taken around the timed region of the benchmarked world itself; a second
    world of the same recipe is then replayed with a hash after EVERY tick (outside every clock).  Both are compared
    with (a) the committed stream tests/golden/stack1m_ref_hashes.npz, written by tests/golden/make_stack1m_ref.py from
    oracle/_ref, and (b) the stream the reference arm of this bench left behind on this box, if it ran here."""
    import numpy as np
    total = args.settle + args.warmup + args.steps
    rec = {"ticks_checked": 0, "equal": None, "against": []}
    streams = []
    if args.entities == 1000000 and os.path.exists(GOLDEN_1M):
        z = np.load(GOLDEN_1M)
        if int(z["cubes"]) == args.entities:
            streams.append(("tests/golden/stack1m_ref_hashes.npz (unmodified reference, %d cubes, committed)" % args.entities,
                            [int(x) for x in z["hashes"]]))
    rf = REF_HASH_FILE % args.entities
    if os.path.exists(rf):
        try:
            j = json.load(open(rf))
            if j.get("cubes") == args.entities and j.get("kind") == "reference":
                streams.append(("reference arm of this bench on this box (%s)" % rf, [int(h, 16) for h in j["hashes"]]))
        except (OSError, ValueError, KeyError):
            pass
    if not streams:
        rec["against"] = ["no reference stream for %d cubes on this box" % args.entities]
        return rec
    w2 = bittoss_b200.World(bittoss_b200.synth_boxstack(args.entities, 4, 777, 64, 60), device=device)
    w2.set_option("rebuild_every", args.rebuild_every)
    mine = [int(x) for x in w2.tick(total, hashes=True)[:, 0]]
    w2.close()
    equal, checked = True, 0
    for name, ref in streams:
        k = min(len(ref), total)
        ok = mine[:k] == ref[:k]
        cp_ok = all(t - 1 >= len(ref) or ref[t - 1] == h for t, h in checkpoints.items() if t >= 1)
        rec["against"].append({"stream": name, "ticks": k, "equal": bool(ok), "timed_world_checkpoints_equal": bool(cp_ok),
                               "first_difference": None if ok else next(i for i in range(k) if mine[i] != ref[i])})
        equal = equal and ok and cp_ok
        checked = max(checked, k)
    rec["ticks_checked"], rec["equal"] = checked, bool(equal)
    rec["checkpoints"] = {str(t): "%016x" % h for t, h in sorted(checkpoints.items())}
    return rec


GOLDEN_1M_PUSHALL = os.path.join(ROOT, "tests", "golden", "stack1m_pushall_ref_hashes.npz")
ALG_BYTES = {  # ALGORITHMIC bytes per unit of each kernel class (DESIGN.md §3 states the per-unit figures)
    "begin_integrate": 94, "decide": 10, "post": 9, "flush1": 5, "hash": 61, "collide": COLLIDE_BYTES_PER_LEAF,
    "morton_keys": 24, "radix_sort": 24, "lbvh_build": 72, "refit": 212, "push": 200}


def kernel_table(prof, ps, n, ticks, peak):
    """per kernel class: launches, device time (CUDA events around every launch) and achieved algorithmic GB/s"""
    total_ms = sum(v[1] for v in prof.values()) or 1.0
    builds = max(1, ps["pair_builds"])
    P = ps["pairs_cached"] / builds            # cached pairs per full k_pairs run
    sort_passes = 4                            # 8-bit LSD radix passes over 30 + world bits
    work = {  # bytes moved by ALL launches of the class in the profiled ticks
        "narrow": 8 * P * prof.get("narrow", (0, 0))[0] + 64 * ps["candidate_pairs"],
        "pairs": (32 * n + 8 * P) * ps["pair_builds"], "refit": 212 * n * ps["pair_builds"],
        "push": 200 * ps["pushes"], "morton_keys": 24 * n * ps["tree_builds"],
        "radix_sort": 24 * n * sort_passes * ps["tree_builds"], "lbvh_build": 72 * n * ps["tree_builds"]}
    out = {}
    for k, (launches, kms) in prof.items():
        if not launches:
            continue
        b = work[k] if k in work else (ALG_BYTES[k] * n * launches if k in ALG_BYTES else None)
        ach = b / (kms * 1e-3) / 1e9 if b and kms > 0 else None
        out[k] = {"launches_per_tick": launches / ticks, "avg_launch_ms": round(kms / launches, 5),
                  "ms_per_tick": round(kms / ticks, 5), "share_of_step": round(kms / total_ms, 4),
                  "alg_gbs": round(ach, 2) if ach else None, "frac": round(ach / peak, 5) if ach else None}
    return out


def dynamic_phase(bittoss_b200, args, device, peak):
    """The regime in which the broadphase runs: the same world with the top cube of EVERY column kicked (250 000
    sliding cubes, 4 067 crushed): first ticks from the upload, so the Morton sort, the LBVH build, refits, full
    pair builds and patches are all inside.  Device-timed unprofiled, then a second world with events around every
    launch for the per-kernel table; the per-tick hashes of a third are compared with the reference's stream."""
    import numpy as np
    ticks = args.dynamic_ticks
    mk = lambda: bittoss_b200.World(bittoss_b200.synth_boxstack(args.entities, 4, 777, 1, 60), device=device)
    w = mk()
    n = w.count()[0]
    ms = w.tick_timed(ticks)
    st = w.stats()
    w.close()
    w = mk()
    w.profile_enable(True)
    w.tick(ticks)
    prof, ps = w.profile(), w.stats()
    w.close()
    rec = {"workload": "the box-stack world with the top cube of every column kicked (push_every=1, |v| <= 60), "
                       "first %d ticks from the upload" % ticks,
           "ticks": ticks, "ms_per_step": ms / ticks, "value": st["entity_ticks"] / (ms * 1e-3),
           "iterations_per_tick": st["iterations"] / ticks, "tree_builds": st["tree_builds"],
           "pair_builds": st["pair_builds"], "pair_patches": st["pair_patches"], "crushes": st["crushes"],
           "pairs_cached": st["pairs_cached"], "per_kernel": kernel_table(prof, ps, n, ticks, peak), "parity": None}
    if args.entities == 1000000 and os.path.exists(GOLDEN_1M_PUSHALL) and not args.no_parity:
        z = np.load(GOLDEN_1M_PUSHALL)
        ref = [int(x) for x in z["hashes"]]
        k = min(len(ref), ticks)
        w = mk()
        mine = [int(x) for x in w.tick(k, hashes=True)[:, 0]]
        w.close()
        rec["parity"] = {"ticks_checked": k, "equal": mine == ref[:k],
                         "against": "tests/golden/stack1m_pushall_ref_hashes.npz (unmodified reference, committed)"}
    return rec


def raycast_record(bittoss_b200, w, args, peak):
    """BASELINE.json configs[4]: 4 M rays against the settled 1 M-entity world `w`; a subsample is checked against the
    reference's own selectCast (raycast.cpp:16-57) on a gamestate rebuilt from the downloaded device state."""
    import numpy as np
    from bittoss_b200 import workloads
    ms, winners, origin, look, build_ms = workloads.raycast_sweep(w, args.rays)
    t0 = time.perf_counter()
    host_out = w.raycast(origin, look)
    host_s = time.perf_counter() - t0
    rays_s = args.rays / (ms * 1e-3)
    rec = {"metric": "rays_per_sec", "value": rays_s, "unit": "rays/s", "rays": args.rays, "entities": w.count()[0],
           "ms": ms, "tree_build_ms": build_ms, "hit_fraction": float((winners >= 0).mean()),
           "host_buffer_rays_per_sec": args.rays / host_s, "host_buffer_equals_device": bool(np.array_equal(winners, host_out)),
           "roofline": {"bound": "hbm", "alg_bytes_per_ray": 40, "achieved_gbs": rays_s * 40 / 1e9, "peak": peak,
                        "frac": rays_s * 40 / 1e9 / peak}}
    if args.ray_check > 0:
        from oracle import pybind
        state = w.download()
        k = min(args.ray_check, args.rays)
        if pybind.have_ref():
            with pybind.stdout_to_stderr():
                R = pybind.RefLib()
                rw = R.new_world()
                R.build_from(state)
                t0 = time.perf_counter()
                want = R.selectcast(origin[:k], look[:k])
                cpu_s = time.perf_counter() - t0
                R.free_world(rw)
            kind = "reference"
        else:
            ow = pybind.OracleLib().create(state)
            t0 = time.perf_counter()
            want = ow.selectcast(origin[:k], look[:k])
            cpu_s = time.perf_counter() - t0
            kind = "port"
        rec["parity_subsample"] = k
        rec["parity_ok"] = bool(np.array_equal(want, winners[:k]))
        rec["cpu_baseline"] = {"kind": kind, "cores": 1, "rays_per_sec": k / cpu_s,
                               "sample": "%d rays, selectCast brute force over %d entities" % (k, state.n)}
    return rec


def arena_record(bittoss_b200, args, device, rank, world_size):
    """BASELINE.json configs[3]: `--arena-worlds` independent arena.sav worlds split over the ranks in contiguous
    blocks (parallel.shard_range), no collective on the tick path, the per-world state hashes all-gathered at the end;
    rank 0 then runs ALL the worlds in one process and checks every gathered hash against that run."""
    from bittoss_b200 import workloads
    from bittoss_b200.parallel import gather_hashes, reduce_time_and_work, shard_range
    level = os.path.join(ROOT, "host", "_build", "data", "arena.sav")
    if not os.path.exists(level):
        level = os.path.join(ROOT, "oracle", "_ref", "data", "arena.sav")
    if not os.path.exists(level):
        return {"skipped": "arena.sav not shipped (host/_build/data is built from the reference tree)"}
    b, e = shard_range(args.arena_worlds, rank, world_size)
    # CUDA loads a kernel's module at its first launch (k_tick_small is the largest of the library): paid on a throw-away batch
    for n_warm in (2, min(e - b, 2400)):   # both variants of the kernel: few worlds / enough worlds to fill the SMs
        warm, _ = workloads.arena_batch(level, list(range(n_warm)))
        ww = workloads.arena_world(warm, device)
        ww.tick(2)
        ww.close()
    batch, n1 = workloads.arena_batch(level, list(range(b, e)))
    w = workloads.arena_world(batch, device)
    ms = w.tick_timed(args.arena_ticks)
    st = w.stats()
    hashes = w.hash()
    w.close()
    dev = "cuda" if world_size > 1 else None
    ms_max, et = reduce_time_and_work(ms, float(st["entity_ticks"]), device=dev)
    all_hashes = gather_hashes(hashes, args.arena_worlds, device=dev)
    rec = None
    if rank == 0:
        rec = {"workload": "%d independent arena.sav worlds (BASELINE.json configs[3]), each with its own seeded kick, "
                           "contiguous shards over %d GPU(s), %d ticks" % (args.arena_worlds, world_size, args.arena_ticks),
               "metric": "world_ticks_per_sec", "value": args.arena_worlds * args.arena_ticks / (ms_max * 1e-3),
               "unit": "world-ticks/s", "entity_ticks_per_sec": et / (ms_max * 1e-3), "n_gpus": world_size,
               "worlds": args.arena_worlds, "worlds_per_gpu": e - b, "entities_per_world": n1, "ticks": args.arena_ticks,
               "ms_per_tick": ms_max / args.arena_ticks, "iterations_per_tick": st["iterations"] / max(1, st["ticks"]),
               "kernel_launches_per_tick": st["kernel_launches"] / max(1, st["ticks"]),
               "schedule": "one launch per tick, one thread block per world (k_tick_small: loop control, candidate pairs and "
                           "the serial lane of a world in shared memory)",
               "unsupported_events": st["unsupported_events"], "distinct_hashes": len(set(all_hashes)),
               "hash_gather": "all_gather of 8 B per world (NCCL)" if world_size > 1 else "single process"}
        if world_size > 1:
            full, _ = workloads.arena_batch(level, list(range(args.arena_worlds)))
            w1 = workloads.arena_world(full, device)
            w1.tick(args.arena_ticks)
            alone = w1.hash()
            w1.close()
            rec["sharded_equals_single_process"] = bool([int(x) for x in alone] == [int(x) for x in all_hashes])
        else:
            k = min(4, args.arena_worlds)   # the same worlds alone must give the same hashes
            ok = True
            for g in range(k):
                one, _ = workloads.arena_batch(level, [g])
                w1 = workloads.arena_world(one, device)
                w1.tick(args.arena_ticks)
                ok = ok and int(w1.hash()[0]) == int(all_hashes[g])
                w1.close()
            rec["alone_equals_batched"] = {"worlds_checked": k, "equal": bool(ok)}
        # the generic schedule (~30 launches per tick over all worlds, pair lists cached across ticks) on the first worlds
        k = min(512, e - b)
        part, _ = workloads.arena_batch(level, list(range(k)))
        w1 = workloads.arena_world(part, device)
        w1.set_option("fused_small", 0)
        g_ms = w1.tick_timed(args.arena_ticks)
        g_st = w1.stats()
        rec["generic_schedule_check"] = {"worlds": k, "equal": bool([int(x) for x in w1.hash()] == [int(x) for x in all_hashes[:k]]),
                                         "kernel_launches_per_tick": g_st["kernel_launches"] / max(1, g_st["ticks"]),
                                         "ms_per_tick": g_ms / args.arena_ticks}
        w1.close()
    return rec


def pin_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs next to its GPU, so that the pinned staging buffers of the end-to-end loop (allocated
    below, first touched here) land on that NUMA node instead of all eight ranks sharing node 0's memory controller."""
    try:
        import torch
        bdf = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        if bdf is None:
            out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
            bdf = out[-12:].lower() if out else None
        else:
            bdf = "%04x:%02x:%02x.0" % (torch.cuda.get_device_properties(local_rank).pci_domain_id,
                                        torch.cuda.get_device_properties(local_rank).pci_bus_id,
                                        torch.cuda.get_device_properties(local_rank).pci_device_id)
        base = "/sys/bus/pci/devices/%s" % bdf
        node = int(open(base + "/numa_node").read().strip())
        cpus = open(base + "/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return {"gpu": local_rank, "pci": bdf, "numa_node": node, "cpus": cpus, "pinned": bool(ids)}
    except Exception as e:  # topology files are best effort: never take the measurement down
        return {"gpu": local_rank, "error": str(e)[:120]}


def run_cuda(args):
    import numpy as np
    import torch
    import bittoss_b200

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world_size > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    numa = pin_to_gpu_numa_node(local_rank) if world_size > 1 else None
    # a throw-away small world first: pays CUDA's lazy module loading outside every measured phase
    _warm = bittoss_b200.World(bittoss_b200.synth_boxstack(4096, 4, 1, 64, 60), device=local_rank)
    _warm.tick(3)
    _warm.close()
    # one independent world per GPU (a single world does not shard)
    wa = bittoss_b200.synth_boxstack(args.entities, 4, 777 + rank, 64, 60)
    n = wa.n
    w = bittoss_b200.World(wa, device=local_rank)
    w.set_option("rebuild_every", args.rebuild_every)
    for kv in args.opt:  # experiments only: any bt_world_set_option knob; none of them changes results
        k, v = kv.split("=")
        w.set_option(k, int(v))

    # ---- (0) settle phase ("run 100 ticks after a 50-tick settle; report
    #      both phases"): the cubes start 100-164 units apart and drop onto each other ----
    settle_ms = w.tick_timed(args.settle) if args.settle > 0 else 0.0
    settle_stats = w.stats(reset=True)
    checkpoints = {args.settle: int(w.hash()[0])}
    # ---- (1) device-resident throughput: `value` ----
    sampler = ClockSampler(local_rank)
    sampler.start()  # ahead of the warm-up, so that nvidia-smi is already reporting when the timed region begins
    w.tick(args.warmup)
    w.stats(reset=True)
    checkpoints[args.settle + args.warmup] = int(w.hash()[0])
    barrier()
    sampler.mark()
    ms = w.tick_timed(args.steps)
    barrier()
    st = w.stats()
    h_final = w.hash()[0]
    checkpoints[args.settle + args.warmup + args.steps] = int(h_final)
    # K ticks of 0.8 ms are over before nvidia-smi's next sample: when the timed region was shorter than
    # 0.6 s the SAME ticks keep running, untimed, right behind it until the sampler has 8 readings under that
    # load (3 s at most; value / ms_per_step are already fixed; the state hash above was taken before)
    in_region = sampler.since_mark()
    clock_window = "timed region"
    try:
        if ms < 600.0 and sampler.p:
            per_tick = max(ms / max(1, args.steps), 1e-3)
            chunk = int(min(1000, max(1, 100.0 / per_tick)))  # about 100 ms of ticks at a time
            extra, t_end = 0, time.perf_counter() + 3.0
            while sampler.since_mark() < 8 and time.perf_counter() < t_end:
                w.tick(chunk)
                w.sync()
                extra += chunk
            clock_window = "timed region (%.0f ms) + %d untimed ticks of the same world right behind it" % (ms, extra)
    except Exception as e:  # never let the clock evidence take the measurement down
        clock_window = "timed region (extension failed: %s)" % e
    clocks = sampler.stop()
    clocks["samples_in_timed_region"] = in_region
    clocks["window"] = clock_window
    from bittoss_b200.parallel import gather_hashes, reduce_time_and_work, shard_range
    assert shard_range(world_size, rank, world_size) == (rank, rank + 1)  # one whole world per rank
    ms_max, et_sum = reduce_time_and_work(ms, float(st["entity_ticks"]), device="cuda")
    # the only collective of the run: the final per-world state hashes
    hashes = gather_hashes([h_final], world_size, device="cuda")
    value = et_sum / (ms_max * 1e-3)

    # ---- (2) end to end through the C ABI with HOST buffers: `e2e` ----
    # per step: H2D of the host-side deltas (uCenter/uVel of the tick handlers, 24 B/entity), the tick, D2H of centre +
    # velocity (24 B/entity) and of the state hash.  Two forms of the same exchange, both timed:
    #   serial     apply_host_deltas -> step -> read_state -> hash, everything on the world's stream, one after the other
    #              (a host whose next deltas need the state it has just read: round 1's `e2e`)
    #   pipelined  the *_async calls: the copies on their own streams, double-buffered, so that H2D of step t+1 and D2H of
    #              step t-1 overlap the tick of step t (a host whose handlers are input-driven, a renderer one tick behind);
    #              every step still copies its own 24 + 24 B/entity inside the timed region and reads its own hash
    live = w.count()[0]
    bufs = [[torch.zeros((live, 3), dtype=torch.int32).pin_memory().numpy() for _ in range(4)] for _ in range(2)]
    e2e_steps = max(1, min(args.steps, 20))

    def e2e_serial(n_steps):
        nonlocal live
        dc, dv, oc, ov = bufs[0]
        et = 0
        for _ in range(n_steps):
            w.apply_host_deltas(dc[:live], dv[:live])
            w.step(31)
            live = w.live_ents()   # host-side bookkeeping of the step that just finished: no extra round trip
            w.read_state(oc[:live], ov[:live])
            _ = w.hash()
            et += live
        return et

    def e2e_pipelined(n_steps):
        nonlocal live
        et = 0
        tickets = [None, None]
        w.apply_host_deltas_async(bufs[0][0][:live], bufs[0][1][:live])
        for t in range(n_steps):
            k = t & 1
            if t + 1 < n_steps:   # the next step's deltas start travelling now: bt_world_step waits for its own tick
                w.send_host_deltas(bufs[k ^ 1][0][:live], bufs[k ^ 1][1][:live])
            w.step(31)
            live_now = w.live_ents()
            _ = w.hash()          # this step's result (12 bytes, through mapped host memory: not behind the state copies)
            tickets[k] = w.read_state_async(bufs[k][2][:live_now], bufs[k][3][:live_now])   # travels during the next tick
            if t + 1 < n_steps:
                if live_now == live:
                    w.commit_host_deltas()   # ... and are added in front of the next step
                else:   # an entity died: the rows sent no longer line up with the ranks; send this step's again, serially
                    try:
                        w.commit_host_deltas()
                    except Exception:
                        pass
                    w.apply_host_deltas_async(bufs[k ^ 1][0][:live_now], bufs[k ^ 1][1][:live_now])
            if tickets[k ^ 1] is not None:
                w.read_state_wait(tickets[k ^ 1])   # the state of the step before is on the host by now
                tickets[k ^ 1] = None
            live = live_now
            et += live
        for tk in tickets:
            if tk is not None:
                w.read_state_wait(tk)
        w.sync()
        return et

    e2e_serial(2)
    e2e_pipelined(2)   # both paths warm (buffers, streams)
    barrier()
    t0 = time.perf_counter()
    et_ser = e2e_serial(e2e_steps)
    barrier()
    ser_s = time.perf_counter() - t0
    ser_max, ser_work = reduce_time_and_work(ser_s, float(et_ser), device="cuda")
    barrier()
    t0 = time.perf_counter()
    et_e2e = e2e_pipelined(e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_max, e2e_work = reduce_time_and_work(e2e_s, float(et_e2e), device="cuda")
    e2e_value = e2e_work / e2e_max
    e2e_serial_value = ser_work / ser_max

    # ---- (3) per-kernel device times (separate pass, CUDA events around every launch) ----
    w.profile_enable(True)
    w.profile(reset=True)
    w.stats(reset=True)
    prof_ticks = 8
    w.tick(prof_ticks)
    prof = w.profile()
    ps = w.stats()
    w.profile_enable(False)
    peak, peak_src = peaks()
    per_kernel = kernel_table(prof, ps, n, prof_ticks, peak)
    name = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_tick"])
    top = per_kernel[name]
    traffic = traffic_src = None
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("kernel") == name:
            traffic = tj.get("dram_bytes_per_launch")
            traffic_src = "%s; %s" % (tj.get("source"), tj.get("measured_at_commit"))
    roofline = {"bound": "hbm", "kernel": "k_" + name, "achieved": top["alg_gbs"], "peak": peak, "unit": "GB/s",
                "frac": top["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "avg_launch_ms": top["avg_launch_ms"], "share_of_step": top["share_of_step"],
                "whole_tick_algorithmic_gbs": value * BYTES_PER_ENTITY_TICK / 1e9,
                "whole_tick_frac": value * BYTES_PER_ENTITY_TICK / 1e9 / peak,
                "per_kernel": per_kernel}

    # ---- (4) the other two BASELINE workloads and the regime in which the broadphase runs ----
    rays = dyn = None
    if rank == 0 and world_size == 1 and not args.no_extras:
        rays = raycast_record(bittoss_b200, w, args, peak)
    w.close()
    if rank == 0 and world_size == 1 and not args.no_extras:
        dyn = dynamic_phase(bittoss_b200, args, local_rank, peak)
    arena = None if args.no_extras else arena_record(bittoss_b200, args, local_rank, rank, world_size)
    if rank == 0:
        cb = None
        if world_size == 1 and not args.no_cpu_baseline:
            # bounded sample of the same workload (same recipe, settle included), sized for ~30 s
            cb = cpu_baseline(min(args.entities, 100000), 1, args.cpu_ticks, args.settle)
        parity = parity_record(bittoss_b200, args, local_rank, checkpoints) if not args.no_parity else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": "synthetic 1M-entity integer AABB box-stack world with gravity and pushes "
                                       "(BASELINE.json configs[2]), one independent world per GPU",
                           "entities_per_gpu": n, "cubes": args.entities, "height": 4, "push_every": 64,
                           "push_max": 60, "rebuild_every": args.rebuild_every,
                           "settle_ticks": args.settle,
                           "settle_phase": {"ms_per_step": settle_ms / max(1, args.settle),
                                            "value": settle_stats["entity_ticks"] / max(1e-9, settle_ms * 1e-3),
                                            "iterations_per_tick": settle_stats["iterations"] / max(1, args.settle),
                                            "pair_builds_per_tick": settle_stats["pair_builds"] / max(1, args.settle)},
                           "l2": "state arrays (>=330 MB/world) exceed the 126 MB L2; no explicit flush",
                           "iterations_per_tick": st["iterations"] / max(1, st["ticks"]),
                           "pushes_per_tick": st["pushes"] / max(1, st["ticks"]),
                           "final_state_hashes": ["%016x" % int(x) for x in hashes],
                           "pairs_cached_per_tick": st["pairs_cached"] / max(1, st["ticks"]),
                           "pair_builds_per_tick": st["pair_builds"] / max(1, st["ticks"])},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 24 * live,
                        "d2h_bytes_per_step": 24 * live + 8, "steps": e2e_steps,
                        "exchange": "pipelined (bt_world_*_async): every step copies its deltas in (pinned, H2D) and its centre + "
                                    "velocity and hash out (D2H) inside the timed region; the copies of neighbouring steps run on their "
                                    "own streams and overlap the tick in between",
                        "serial_value": e2e_serial_value,
                        "serial_exchange": "H2D deltas -> tick -> D2H centre + velocity -> D2H hash, one after the other on one stream "
                                           "(round 1's e2e)",
                        "rank0_numa": numa},
                "gpu_launches": st["kernel_launches"],
                "roofline": roofline,
                "parity": parity,
                "dynamic_phase": dyn,
                "raycast": rays,
                "batched_arena": arena,
                "cpu_baseline": cb}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else any library prints on
    fd 1 (NCCL's version banner, the reference's puts()) was pointed at stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--entities", type=int, default=1000000)
    ap.add_argument("--rebuild-every", type=int, default=16)
    ap.add_argument("--cpu-ticks", type=int, default=10)
    ap.add_argument("--settle", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the replay against the reference hash streams")
    ap.add_argument("--no-extras", action="store_true", help="skip the dynamic phase, the raycast sweep and the batched arena worlds")
    ap.add_argument("--dynamic-ticks", type=int, default=12)
    ap.add_argument("--rays", type=int, default=4000000)
    ap.add_argument("--ray-check", type=int, default=256, help="rays checked against the reference's selectCast (brute force on the CPU)")
    ap.add_argument("--arena-worlds", type=int, default=8192)
    ap.add_argument("--arena-ticks", type=int, default=200)
    ap.add_argument("--opt", action="append", default=[], help="name=value for bt_world_set_option (tuning experiments)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
