"""BASELINE.json configs[4]: batched raycast sweep — 4 M rays against the 1 M-entity world
(the box-stack world after a settle), SURVEY.md §8(d) "Config 5".

    python tools/raycast_bench.py [--entities 1000000] [--rays 4000000] [--settle 50] [--check 4096]

Prints one JSON line: rays/s with rays and results resident in HBM (CUDA events on the
world's stream), the same through the host-buffer C ABI call, parity of a subsample against
the oracle's brute-force selectCast, and the CPU brute-force rate (extrapolated, stated)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def splitmix32_first(seed):
    """first output of splitmix32 (random.cpp:18-24) for an array of seeds"""
    z = (seed.astype(np.uint64) + np.uint64(0x9e3779b9)) & np.uint64(0xffffffff)
    z ^= z >> np.uint64(16); z = (z * np.uint64(0x21f0aaad)) & np.uint64(0xffffffff)
    z ^= z >> np.uint64(15); z = (z * np.uint64(0x735a2d97)) & np.uint64(0xffffffff)
    z ^= z >> np.uint64(15)
    return z


def make_rays(n_rays, n_ents):
    i = np.arange(n_rays, dtype=np.uint64)
    origin = ((i * np.uint64(2654435761)) & np.uint64(0xffffffff)) % np.uint64(n_ents)
    look = np.zeros((n_rays, 3), dtype=np.int32)
    for k in range(3):  # three consecutive draws of the stream seeded 4242+i
        s = (np.uint64(4242) + i + np.uint64(k) * np.uint64(0x9e3779b9)) & np.uint64(0xffffffff)
        look[:, k] = (splitmix32_first(s) % np.uint64(129)).astype(np.int32) - 64
    look[np.abs(look).sum(1) == 0] = (0, 0, 64)
    return origin.astype(np.int32), look


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--entities", type=int, default=1000000)
    ap.add_argument("--rays", type=int, default=4000000)
    ap.add_argument("--settle", type=int, default=50)
    ap.add_argument("--check", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import bittoss_b200
    wa = bittoss_b200.synth_boxstack(args.entities, 4, 777, 64, 60)
    w = bittoss_b200.World(wa)
    w.tick(args.settle)
    live = w.count()[0]
    origin, look = make_rays(args.rays, live)
    d_o = torch.from_numpy(origin).cuda()
    d_l = torch.from_numpy(look).cuda()
    d_w = torch.empty(args.rays, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    w.raycast_build()
    w.sync()
    stream = torch.cuda.ExternalStream(w.stream())
    times = []
    for _ in range(args.reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            w.raycast_device(d_o.data_ptr(), d_l.data_ptr(), args.rays, d_w.data_ptr())
            e1.record(stream)
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times[1:])
    t0 = time.perf_counter()
    host_out = w.raycast(origin, look)
    host_s = time.perf_counter() - t0
    winners = d_w.cpu().numpy()
    assert np.array_equal(winners, host_out)
    # parity on a subsample against the oracle's brute force over the downloaded state
    from oracle.pybind import OracleLib
    state = w.download()
    ow = OracleLib().create(state)
    k = min(args.check, args.rays)
    t0 = time.perf_counter()
    want = ow.selectcast(origin[:k], look[:k])
    cpu_s = time.perf_counter() - t0
    ok = bool(np.array_equal(want, winners[:k]))
    peak = 6550.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    rays_s = args.rays / (ms * 1e-3)
    print(json.dumps({
        "metric": "rays_per_sec", "value": rays_s, "unit": "rays/s", "rays": args.rays, "entities": live,
        "ms": ms, "hit_fraction": float((winners >= 0).mean()),
        "parity_subsample": k, "parity_ok": ok,
        "host_buffer_rays_per_sec": args.rays / host_s,
        "roofline": {"bound": "hbm", "alg_bytes_per_ray": 40, "achieved_gbs": rays_s * 40 / 1e9,
                     "peak": peak, "frac": rays_s * 40 / 1e9 / peak},
        "cpu_baseline": {"kind": "port", "cores": 1, "rays_per_sec": k / cpu_s,
                         "sample": "%d rays brute force over %d entities (oracle selectCast), "
                                   "rate extrapolates linearly in rays" % (k, live)}}))
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
