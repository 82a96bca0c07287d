"""Per-kernel-class device time of the first ticks of the bench world (stacks settling: crushes,
sliding cubes, pair-cache patches / rebuilds).  Usage: python tools/settle_profile.py [cubes] [ticks] [name=value ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bittoss_b200

cubes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 50
w = bittoss_b200.World(bittoss_b200.synth_boxstack(cubes, 4, 777, 64, 60))
w.tick(3)   # pays CUDA's lazy module loading once, outside the measurement
w.close()
w = bittoss_b200.World(bittoss_b200.synth_boxstack(cubes, 4, 777, 64, 60))
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    w.set_option(k, int(v))
ms = w.tick_timed(ticks)
st = w.stats()
print("unprofiled: %.4f ms/tick, iterations/tick %.2f, pair builds %d, patches %d, tree builds %d, crushes %d" % (
    ms / ticks, st["iterations"] / ticks, st["pair_builds"], st["pair_patches"], st["tree_builds"], st["crushes"]))
w.close()
w = bittoss_b200.World(bittoss_b200.synth_boxstack(cubes, 4, 777, 64, 60))
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    w.set_option(k, int(v))
w.profile_enable(True)
w.tick(ticks)
p = w.profile()
for name, (launches, t) in sorted(p.items(), key=lambda kv: -kv[1][1]):
    if launches:
        print("   %-16s %8.4f ms/tick  %6.2f launches/tick  avg %.4f ms" % (name, t / ticks, launches / ticks, t / launches))
w.close()
