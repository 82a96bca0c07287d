"""Generates tests/golden/tf_*.npz by RUNNING THE REFERENCE with the full game (players with
seeded inputs, tick handlers, world rules — oracle/ref_harness.cpp btref_pre_physics mode 0):
for every tick the exported state right before doPhysics, and the reference's state hash
after its own doPhysics + finishStep.  Lets machines without the reference tree run the
teacher-forced parity check of tests/test_teacher_forced.py.

    python tests/golden/make_teacher_forced.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from util import GOLDEN  # noqa: E402
from oracle.pybind import RefLib  # noqa: E402

CASES = {"tf_arena_2players": ("arena", (1, 2), 0, 160), "tf_lobby_4players": ("lobby", (0, 0, 0, 0), 0, 60),
         "tf_spacetrash_spawner": ("spacetrash", (), 1, 120)}


def make(name):
    level, players, st_on, ticks = CASES[name]
    R = RefLib()
    wid = R.new_world()
    R.load_sav(level + ".sav", players=players)
    R.set_options(st_on, 0)
    fields = None
    chunks, sizes, hashes = {}, [], []
    for t in range(ticks):
        R.pre_physics(0)
        wa = R.export()
        R.do_physics()
        hashes.append(R.hash())
        n, ns, nw = wa.n, int(wa.desc.n_sliders), int(wa.desc.n_wires)
        sizes.append((n, ns, nw))
        for k, v in wa.a.items():
            chunks.setdefault(k, []).append(v.copy())
    R.free_world(wid)
    out = {"sizes": np.array(sizes, dtype=np.int64), "ref_hashes": np.array(hashes, dtype=np.uint64)}
    for k, lst in chunks.items():
        out["f_" + k] = np.concatenate(lst)
        out["o_" + k] = np.cumsum([0] + [len(x) for x in lst]).astype(np.int64)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "ticks", ticks, "ents", sizes[0][0], "->", sizes[-1][0], "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        make(nm)
