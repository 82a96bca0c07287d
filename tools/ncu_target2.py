"""Fixed workloads for ncu (run with `--profile-from-start off`: only the bracketed part is captured).
    python tools/ncu_target2.py dynamic [ticks=3]   the push-all 1M world from its upload: Morton keys, radix sort, LBVH
                                                    build, full refit + pair build, patches, the resolve loop
    python tools/ncu_target2.py steady [ticks=1]    the settled bench world (56 ticks in)
    python tools/ncu_target2.py raycast [rays=4000000]  ray-tree build + one sweep over the settled world"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bittoss_b200

mode = sys.argv[1] if len(sys.argv) > 1 else "dynamic"
arg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cu = ctypes.CDLL("libcuda.so.1")
# CUDA's lazy module loading is paid on a throw-away world
w0 = bittoss_b200.World(bittoss_b200.synth_boxstack(4096, 4, 1, 64, 60))
w0.tick(3)
w0.raycast([0], [[0, 0, 64]])
w0.close()
if mode == "dynamic":
    w = bittoss_b200.World(bittoss_b200.synth_boxstack(1000000, 4, 777, 1, 60))
    w.sync()
    cu.cuProfilerStart()
    w.tick(arg or 3)
    w.sync()
    cu.cuProfilerStop()
    s = w.stats()
    print("dynamic: ticks", s["ticks"], "launches", s["kernel_launches"], "tree builds", s["tree_builds"], "pair builds",
          s["pair_builds"], "patches", s["pair_patches"], "pairs cached", s["pairs_cached"], "hash", hex(w.hash()[0]))
elif mode == "steady":
    w = bittoss_b200.World(bittoss_b200.synth_boxstack(1000000, 4, 777, 64, 60))
    w.tick(56)
    w.stats(reset=True)
    w.sync()
    cu.cuProfilerStart()
    w.tick(arg or 1)
    w.sync()
    cu.cuProfilerStop()
    s = w.stats()
    print("steady: ticks", s["ticks"], "launches", s["kernel_launches"], "iterations", s["iterations"], "candidate pairs",
          s["candidate_pairs"], "pushes", s["pushes"], "hash", hex(w.hash()[0]))
else:
    from bittoss_b200 import workloads
    w = bittoss_b200.World(bittoss_b200.synth_boxstack(1000000, 4, 777, 64, 60))
    w.tick(50)
    origin, look = workloads.make_rays(arg or 4000000, w.count()[0])
    w.sync()
    cu.cuProfilerStart()
    out = w.raycast(origin, look)
    cu.cuProfilerStop()
    print("raycast: rays", len(out), "hits", int((out >= 0).sum()))
w.close()
