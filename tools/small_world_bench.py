"""ticks/s of small worlds, where per-iteration host syncs dominate: run-ahead on vs off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bittoss_b200
for cubes in (512, 20000, 200000):
    for ra in (0, 1):
        wa = bittoss_b200.synth_boxstack(cubes, 4, 777, 64, 60)
        w = bittoss_b200.World(wa)
        w.set_option("run_ahead", ra)
        w.tick(60)
        ms = w.tick_timed(300)
        st = w.stats()
        print("cubes %7d run_ahead %d: %.1f us/tick, %.1fM entity-ticks/s, hash %016x" % (
            cubes, ra, 1e3 * ms / 300, wa.n * 300 / ms / 1e3, w.hash()[0]))
        w.close()
