"""Short, fixed workload for ncu: the bench world, a few whole ticks (same code path as
bench.py's timed region, nothing else).  Usage: python tools/ncu_target.py [cubes] [ticks] [warm]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bittoss_b200

cubes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 2
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wa = bittoss_b200.synth_boxstack(cubes, 4, 777, 64, 60)
w = bittoss_b200.World(wa)
w.tick(warm)
s0 = w.stats(reset=True)
w.sync()
# under `ncu --profile-from-start off` only the measured ticks are captured
import ctypes
_cu = ctypes.CDLL("libcuda.so.1")
_cu.cuProfilerStart()
w.tick(ticks)
w.sync()
_cu.cuProfilerStop()
s = w.stats()
print("ticks", ticks, "launches", s["kernel_launches"], "iterations", s["iterations"], "hash", hex(w.hash()[0]))
print("warm-up: pair builds", s0["pair_builds"], "pairs cached per build", s0["pairs_cached"] // max(1, s0["pair_builds"]),
      "| measured: candidate pairs per iteration", s["candidate_pairs"] // max(1, s["iterations"]), "positives per iteration",
      s["positives"] // max(1, s["iterations"]))
w.close()
