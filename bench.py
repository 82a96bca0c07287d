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
    # per step: H2D of the host-side deltas (uCenter/uVel of the tick handlers, 24 B/entity),
    # the tick, D2H of centre+velocity (24 B/entity) and of the state hash
    live = w.count()[0]
    dc = torch.zeros((live, 3), dtype=torch.int32).pin_memory().numpy()
    dv = torch.zeros((live, 3), dtype=torch.int32).pin_memory().numpy()
    oc = torch.zeros((live, 3), dtype=torch.int32).pin_memory().numpy()
    ov = torch.zeros((live, 3), dtype=torch.int32).pin_memory().numpy()
    e2e_steps = max(1, min(args.steps, 20))
    barrier()
    t0 = time.perf_counter()
    et_e2e = 0
    for _ in range(e2e_steps):
        w.apply_host_deltas(dc, dv)
        w.step(31)
        live_now = w.stats()["live_ents"]
        if live_now != live:  # an entity died: the host arrays shrink with the world
            live = live_now
            dc, dv, oc, ov = dc[:live], dv[:live], oc[:live], ov[:live]
        w.read_state(oc, ov)
        _ = w.hash()
        et_e2e += live
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_max, e2e_work = reduce_time_and_work(e2e_s, float(et_e2e), device="cuda")
    e2e_value = e2e_work / e2e_max

    # ---- (3) per-kernel device times (separate pass, CUDA events around every launch) ----
    w.profile_enable(True)
    w.profile(reset=True)
    w.stats(reset=True)
    prof_ticks = 8
    w.tick(prof_ticks)
    prof = w.profile()
    ps = w.stats()
    w.profile_enable(False)
    total_ms = sum(v[1] for v in prof.values()) or 1.0
    peak, peak_src = peaks()
    n_leaves = n  # every entity of this world is boxed
    P = ps["pairs_cached"] / max(1, ps["pair_builds"])            # cached pairs per k_pairs launch
    # ALGORITHMIC bytes per launch of each kernel class (DESIGN.md "Kernels" states the per-unit figures)
    alg = {
        "narrow": lambda L: 8 * P + 64 * ps["candidate_pairs"] / max(1, L),
        # k_refit / k_pairs return at once unless the pair cache was invalidated: only the launches
        # that did the work (pair_builds of them) move bytes
        "pairs": lambda L: (32 * n_leaves + 8 * P) * ps["pair_builds"] / max(1, L),
        "collide": lambda L: COLLIDE_BYTES_PER_LEAF * n_leaves,
        "refit": lambda L: 212 * n_leaves * ps["pair_builds"] / max(1, L),
        "push": lambda L: 200 * ps["pushes"] / max(1, L),
        "begin_integrate": lambda L: 94 * n,
        "decide": lambda L: 10 * n,
        "post": lambda L: 9 * n,
        "flush1": lambda L: 5 * n,
        "hash": lambda L: 61 * n,
    }
    per_kernel = {}
    for k, (launches, kms) in prof.items():
        if not launches:
            continue
        avg_ms = kms / launches
        ach = alg[k](launches) / (avg_ms * 1e-3) / 1e9 if k in alg and avg_ms > 0 else None
        if ach is not None and ach <= 0:
            ach = None
        per_kernel[k] = {"launches_per_tick": launches / prof_ticks, "avg_launch_ms": round(avg_ms, 5),
                         "ms_per_tick": round(kms / prof_ticks, 5), "share_of_step": round(kms / total_ms, 4),
                         "alg_gbs": round(ach, 2) if ach else None, "frac": round(ach / peak, 5) if ach else None}
    name = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_tick"])
    top = per_kernel[name]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("kernel") == name:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "k_" + name, "achieved": top["alg_gbs"], "peak": peak, "unit": "GB/s",
                "frac": top["frac"], "traffic": traffic, "peak_source": peak_src,
                "avg_launch_ms": top["avg_launch_ms"], "share_of_step": top["share_of_step"],
                "whole_tick_algorithmic_gbs": value * BYTES_PER_ENTITY_TICK / 1e9,
                "whole_tick_frac": value * BYTES_PER_ENTITY_TICK / 1e9 / peak,
                "per_kernel": per_kernel}

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
                        "d2h_bytes_per_step": 24 * live + 8, "steps": e2e_steps},
                "gpu_launches": st["kernel_launches"],
                "roofline": roofline,
                "parity": parity,
                "cpu_baseline": cb}
        emit(line)
    w.close()
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
