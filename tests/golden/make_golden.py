"""Generates tests/golden/*.npz by RUNNING THE REFERENCE (oracle/_ref/libbtref.so, the
unmodified sboerwinkle/bittoss sources).  Needs /root/reference to have been compiled by
`make -C oracle`; the fixtures are committed so that the GPU box (which has no reference
tree) can check against them.

Each fixture = initial world as plain bt_world_desc arrays + the reference's per-tick state
hash stream ("ref_hashes") + final live count.  Levels are loaded with deserialize(), their
tick handlers dropped and collision-path handlers without a device implementation replaced
by NIL (tests/util.py KEEP), then kicked with seeded velocities through uVel.

    python tests/golden/make_golden.py [name ...]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from util import KEEP, SCENARIOS, golden_path, save_world  # noqa: E402
from oracle.pybind import RefLib  # noqa: E402


def make(name):
    kind, args, ticks = SCENARIOS[name]
    R = RefLib()
    w = R.new_world()
    if kind == "stack":
        R.make_boxstack(*args)
    else:
        lv, seed, maxv = args
        R.load_sav(lv + ".sav")
        R.strip_handlers(True, **KEEP)
        if maxv:
            R.kick(seed, maxv, 1)
    wa = R.export()
    hashes = np.zeros(ticks, dtype=np.uint64)
    for t in range(ticks):
        R.step(1)
        hashes[t] = R.hash()
    n_final = R.count()[0]
    R.free_world(w)
    save_world(golden_path(name), wa, ref_hashes=hashes, n_final=np.array([n_final]))
    print("%-18s ents %6d -> %6d  ticks %3d  last hash %016x" % (name, wa.n, n_final, ticks, int(hashes[-1])))


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(SCENARIOS)):
        make(nm)
