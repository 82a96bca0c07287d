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
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(n_cubes, warm_ticks, ticks):
    """The UNMODIFIED reference's CPU tick (oracle/_ref) when it was compiled, else the C port."""
    from oracle import pybind
    cores = 1  # the reference tick is single-threaded and non-reentrant (ent.cpp:17, velbox/box.cpp:17-18)
    if pybind.have_ref():
        R = pybind.RefLib()
        R.new_world()
        R.make_boxstack(n_cubes, 4, 777, 64, 60)
        for _ in range(warm_ticks):
            R.step(1)
        secs, et = R.time_steps(ticks, 1)
        kind = "reference"
    else:
        import bittoss_b200
        wa = bittoss_b200.synth_boxstack(n_cubes, 4, 777, 64, 60)
        ow = pybind.OracleLib().create(wa)
        for _ in range(warm_ticks):
            ow.step(31)
        secs, et = ow.time_steps(ticks)
        kind = "port"
    return {"value": et / secs, "unit": UNIT, "cores": cores, "kind": kind, "seconds": round(secs, 3),
            "sample": "%d whole ticks of the %d-cube box-stack world after %d warm-up ticks, 1 thread"
                      % (ticks, n_cubes, warm_ticks)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ticks = args.steps + args.warmup
    # bounded sample: keep the whole run within ~2 minutes at ~5.5 us per entity-tick on one core
    n_cubes = int(min(args.entities, max(20000, 110.0 / (ticks * 5.5e-6))))
    cb = cpu_baseline(n_cubes, args.warmup, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cb["seconds"] / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "synthetic box-stack world, gravity + pushes (BASELINE.json configs[2])",
                       "entities": n_cubes, "height": 4, "sample": cb["sample"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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

    # one independent world per GPU (a single world does not shard)
    wa = bittoss_b200.synth_boxstack(args.entities, 4, 777 + rank, 64, 60)
    n = wa.n
    w = bittoss_b200.World(wa, device=local_rank)
    w.set_option("rebuild_every", args.rebuild_every)

    # ---- (1) device-resident throughput: `value` ----
    w.tick(args.warmup)
    w.stats(reset=True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ms = w.tick_timed(args.steps)
    barrier()
    clocks = sampler.stop()
    st = w.stats()
    h_final = w.hash()[0]
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    et = torch.tensor([float(st["entity_ticks"])], dtype=torch.float64, device="cuda")
    hashes = torch.tensor([h_final & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(et, op=dist.ReduceOp.SUM)
        gathered = [torch.zeros_like(hashes) for _ in range(world_size)]
        dist.all_gather(gathered, hashes)  # the only collective: final per-world state hashes
        hashes = torch.cat(gathered)
    ms_max = float(t.item())
    value = float(et.item()) / (ms_max * 1e-3)

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
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    ee = torch.tensor([float(et_e2e)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(ee, op=dist.ReduceOp.SUM)
    e2e_value = float(ee.item()) / float(te.item())

    # ---- (3) per-kernel device times (separate pass, events around every launch) ----
    w.profile_enable(True)
    w.profile(reset=True)
    s0 = w.stats()
    prof_ticks = 4
    w.tick(prof_ticks)
    prof = w.profile()
    s1 = w.stats()
    w.profile_enable(False)
    total_ms = sum(v[1] for v in prof.values()) or 1.0
    top = max(prof.items(), key=lambda kv: kv[1][1])
    peak, peak_src = peaks()
    name, (launches, kms) = top
    n_leaves = n  # every entity of this world is boxed
    achieved = (COLLIDE_BYTES_PER_LEAF * n_leaves) / ((kms / max(1, launches)) * 1e-3) / 1e9 if name == "collide" else None
    roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
                "avg_launch_ms": kms / max(1, launches), "share_of_step": kms / total_ms,
                "whole_tick_algorithmic_gbs": value * BYTES_PER_ENTITY_TICK / 1e9,
                "whole_tick_frac": value * BYTES_PER_ENTITY_TICK / 1e9 / peak,
                "per_kernel_ms_per_tick": {k: round(v[1] / prof_ticks, 4) for k, v in prof.items() if v[0]},
                "per_kernel_launches_per_tick": {k: v[0] / prof_ticks for k, v in prof.items() if v[0]}}

    if rank == 0:
        cb = None
        if world_size == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(args.entities, 1, args.cpu_ticks)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": "synthetic 1M-entity integer AABB box-stack world with gravity and pushes "
                                       "(BASELINE.json configs[2]), one independent world per GPU",
                           "entities_per_gpu": n, "cubes": args.entities, "height": 4, "push_every": 64,
                           "push_max": 60, "rebuild_every": args.rebuild_every,
                           "l2": "state arrays (>=330 MB/world) exceed the 126 MB L2; no explicit flush",
                           "iterations_per_tick": st["iterations"] / max(1, st["ticks"]),
                           "pushes_per_tick": st["pushes"] / max(1, st["ticks"]),
                           "final_state_hashes": [int(x) for x in hashes.tolist()]},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 24 * live,
                        "d2h_bytes_per_step": 24 * live + 8, "steps": e2e_steps},
                "gpu_launches": st["kernel_launches"],
                "roofline": roofline,
                "cpu_baseline": cb}
        print(json.dumps(line), flush=True)
    w.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--entities", type=int, default=1000000)
    ap.add_argument("--rebuild-every", type=int, default=16)
    ap.add_argument("--cpu-ticks", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
