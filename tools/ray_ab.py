"""4 M rays against the settled 1 M-entity world: sweep time per option set (CUDA events), winners compared between them.
    python tools/ray_ab.py "ray_dynamic=0" "ray_dynamic=1" "ray_dynamic=1 sort_rays=1" ..."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bittoss_b200
from bittoss_b200 import workloads

w = bittoss_b200.World(bittoss_b200.synth_boxstack(int(os.environ.get("AB_CUBES", "1000000")), 4, 777, 64, 60))
w.tick(50)
ref = None
tag = os.path.basename(os.environ.get("BT_LIB_SO", "in-tree")).replace("libbt_", "").replace(".so", "")
for spec in sys.argv[1:] or ["ray_dynamic=0", "ray_dynamic=1"]:
    for kv in spec.split():
        k, v = kv.split("=")
        w.set_option(k, int(v))
    ms, winners, _, _, build_ms = workloads.raycast_sweep(w, int(os.environ.get("AB_RAYS", "4000000")))
    if ref is None:
        ref = winners
    print("%-8s %-28s %.3f ms  %.0f M rays/s  same winners as the first: %s" % (
        tag, spec, ms, len(winners) / ms / 1e3, bool(np.array_equal(ref, winners))), flush=True)
w.close()
