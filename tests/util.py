"""Shared helpers for the test-suite: world fixtures on disk, scenario table."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from bittoss_b200.desc import WorldArrays  # noqa: E402

GOLDEN = os.path.join(HERE, "golden")

# device-side handler coverage (bt_phys.cuh BT_SUPPORTED_*): everything else is replaced by
# the catalog's NIL entry in the INPUT world when a shipped level is turned into a fixture
KEEP = dict(keep_push=0b101, keep_pushed=0b101111, keep_pair=0b1001, keep_crush=0b1)

# name -> (kind, args, ticks).  "stack": (n_cubes, height, seed, push_every, push_max)
#                               "level": (file, kick_seed, kick_max_v)
SCENARIOS = {
    "stack512": ("stack", (512, 4, 777, 64, 60), 60),
    "stack4k_pushall": ("stack", (4000, 4, 777, 1, 60), 100),
    "stack2k_8high": ("stack", (2000, 8, 5, 4, 60), 80),
    "stack20k": ("stack", (20000, 4, 99, 64, 60), 40),
    "spider_static": ("level", ("spider", 1, 0), 60),
    "spider_kick": ("level", ("spider", 1000, 200), 40),
    "lobby_kick": ("level", ("lobby", 1000, 300), 60),
    "tower_kick": ("level", ("tower", 5, 300), 60),
    "forest_kick": ("level", ("forest", 5, 300), 60),
    "diner_kick": ("level", ("diner", 9, 200), 60),
    "arena_kick": ("level", ("arena", 3, 300), 80),
    "leggs_kick": ("level", ("leggs", 2, 200), 60),
    "spacetrash_kick": ("level", ("spacetrash", 4, 300), 60),
}


def save_world(path, wa: WorldArrays, **extra):
    n, ns, nw = wa.n, int(wa.desc.n_sliders), int(wa.desc.n_wires)
    out = {"_n": np.array([n, ns, nw, int(wa.desc.n_worlds)], dtype=np.int64)}
    for k, v in wa.a.items():
        out[k] = v
    out.update(extra)
    np.savez_compressed(path, **out)


def load_world(path):
    z = np.load(path)
    n, ns, nw, ng = [int(x) for x in z["_n"]]
    wa = WorldArrays(n, ns, nw, ng)
    for k in wa.a:
        wa.a[k][:] = z[k][: len(wa.a[k])]
    extra = {k: z[k] for k in z.files if k not in wa.a and k != "_n"}
    return wa, extra


def golden_path(name):
    return os.path.join(GOLDEN, name + ".npz")


def load_scenario(name):
    """-> (WorldArrays initial state, reference per-tick hashes uint64[ticks])."""
    wa, extra = load_world(golden_path(name))
    return wa, extra["ref_hashes"]


def load_teacher_forced(name):
    """-> list of (WorldArrays state before doPhysics, reference hash after it), one per tick."""
    z = np.load(golden_path(name))
    sizes, hashes = z["sizes"], z["ref_hashes"]
    out = []
    for t in range(len(sizes)):
        n, ns, nw = [int(x) for x in sizes[t]]
        wa = WorldArrays(n, ns, nw, 1)
        for k in wa.a:
            off = z["o_" + k]
            wa.a[k][:] = z["f_" + k][off[t]: off[t + 1]][: len(wa.a[k])]
        out.append((wa, int(hashes[t])))
    return out
