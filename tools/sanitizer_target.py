"""Small end-to-end run for compute-sanitizer: box stack with crushes, a level with hold trees,
batched worlds, raycasts, host exchange, download."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import bittoss_b200
from util import load_scenario

for name, ticks in (("stack4k_pushall", 40), ("spider_kick", 40), ("diner_kick", 30)):
    wa, ref = load_scenario(name)
    w = bittoss_b200.World(wa)
    w.set_option("rebuild_every", 3)
    hs = w.tick(ticks, hashes=True)
    assert [int(x) for x in hs[:, 0]] == [int(x) for x in ref[:ticks]], name
    rng = np.random.RandomState(1)
    n = w.count()[0]
    origin = rng.randint(0, n, size=500).astype(np.int32)
    look = rng.randint(-64, 65, size=(500, 3)).astype(np.int32)
    look[np.abs(look).sum(1) == 0] = (0, 0, 64)
    w.raycast(origin, look)
    c = np.zeros((n, 3), dtype=np.int32); v = np.zeros((n, 3), dtype=np.int32)
    w.apply_host_deltas(c, v); w.step(31); n = w.count()[0]
    c = np.zeros((n, 3), dtype=np.int32); v = np.zeros((n, 3), dtype=np.int32)
    w.read_state(c, v); w.download(); w.close()
ws = [load_scenario(n)[0] for n in ("arena_kick", "stack512", "tower_kick")]
w = bittoss_b200.World(bittoss_b200.concat_worlds(ws)); w.tick(30); w.hash(); w.close()
w = bittoss_b200.World(bittoss_b200.synth_boxstack(3000)); w.set_option("pair_cache", 0); w.tick(10); w.close()
print("sanitizer target ok")
