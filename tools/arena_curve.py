"""Tick time of batched arena.sav worlds (BASELINE.json configs[3] shape) against the number of worlds in one device world:
what the fixed cost per tick is, and what a GPU does with 1/8 of the 8192 worlds.  python tools/arena_curve.py [opt=value ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bittoss_b200
from bittoss_b200 import workloads

level = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "host", "_build", "data", "arena.sav")
opts = [kv.split("=") for kv in sys.argv[1:] if not kv.startswith("only=")]
sizes = [int(kv[5:]) for kv in sys.argv[1:] if kv.startswith("only=")] or [256, 1024, 4096, 8192]
tag = os.path.basename(os.environ.get("BT_LIB_SO", "in-tree")).replace("libbt_", "").replace(".so", "")
for nw in sizes:
    batch, n1 = workloads.arena_batch(level, list(range(nw)))
    w = workloads.arena_world(batch, 0)
    for k, v in opts:
        w.set_option(k, int(v))
    first = w.tick_timed(200) / 200          # from the upload: every entity kicked, the busy phase
    ms = min(w.tick_timed(200) for _ in range(2)) / 200
    st = w.stats()
    h = w.hash()
    print("%-8s %s worlds %5d  first 200 ticks %.4f ms/tick, then %.4f ms/tick  %6.1f M world-ticks/s  iters/tick %.2f  launches/tick %.1f  hash0 %016x" % (
        tag, " ".join("%s=%s" % (k, v) for k, v in opts), nw, first, ms, nw / ms / 1e3, st["iterations"] / max(1, st["ticks"]),
        st["kernel_launches"] / max(1, st["ticks"]), h[0] ^ h[-1]), flush=True)
    w.close()
