"""One line per library build: ms/tick of the settled 1M-cube world (device-timed, 60 ticks after a 50-tick settle + 5
warm-up), the settle phase, per-kernel-class times from a profiled pass, and the state hash.  The library is chosen
with BT_LIB_SO (tools/ab.sh loops over build/libbt_<name>.so variants on ONE box: boxes of the pool differ by several %)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bittoss_b200

cubes = int(os.environ.get("AB_CUBES", "1000000"))
push_every = int(os.environ.get("AB_PUSH_EVERY", "64"))
opts = [kv.split("=") for kv in sys.argv[1:]]
w = bittoss_b200.World(bittoss_b200.synth_boxstack(4096, 4, 1, 64, 60))
w.tick(3)
w.close()
if os.environ.get("AB_DYNAMIC"):   # the broadphase regime: first 12 ticks of the push-all world, per-kernel profile of the same
    wa = bittoss_b200.synth_boxstack(cubes, 4, 777, 1, 60)
    w = bittoss_b200.World(wa)
    for k, v in opts:
        w.set_option(k, int(v))
    ms = w.tick_timed(12) / 12
    h = w.hash()[0]
    w.close()
    w = bittoss_b200.World(wa)
    for k, v in opts:
        w.set_option(k, int(v))
    w.profile_enable(True)
    w.profile(reset=True)
    w.tick(12)
    p = w.profile()
    w.close()
    per = " ".join("%s %.1f" % (k, 1e3 * t / 12) for k, (n, t) in sorted(p.items(), key=lambda kv: -kv[1][1]) if n and t / 12 > 0.004)
    print("%-10s dynamic %.4f ms/tick  hash %016x | us/tick: %s" % (
        os.path.basename(os.environ.get("BT_LIB_SO", "in-tree")).replace("libbt_", "").replace(".so", ""), ms, h, per), flush=True)
    sys.exit(0)
w = bittoss_b200.World(bittoss_b200.synth_boxstack(cubes, 4, 777, push_every, 60))
for k, v in opts:
    w.set_option(k, int(v))
settle = w.tick_timed(50) / 50
w.tick(5)
w.stats(reset=True)
ms = min(w.tick_timed(60) / 60 for _ in range(3))
st = w.stats()
h = w.hash()[0]
w.profile_enable(True)
w.profile(reset=True)
w.tick(8)
p = w.profile()
w.profile_enable(False)
w.close()
per = " ".join("%s %.1f" % (k, 1e3 * t / 8) for k, (n, t) in sorted(p.items(), key=lambda kv: -kv[1][1]) if n and t / 8 > 0.004)
print("%-10s %.4f ms/tick  settle %.4f  iters/tick %.2f  hash %016x | us/tick: %s" % (
    os.path.basename(os.environ.get("BT_LIB_SO", "in-tree")).replace("libbt_", "").replace(".so", ""), ms, settle,
    st["iterations"] / max(1, st["ticks"]), h, per), flush=True)
