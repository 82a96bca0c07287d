"""bt_world_dup / bt_world_assign (dup(gs), ent.cpp:1135-1200: the client's rollback snapshot) on the bench world."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bittoss_b200

cubes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
w = bittoss_b200.World(bittoss_b200.synth_boxstack(cubes, 4, 777, 64, 60))
w.tick(20)
t0 = time.perf_counter(); snap = w.dup(); snap.sync(); t_first = time.perf_counter() - t0
ts = []
for i in range(10):
    w.tick(2)
    t0 = time.perf_counter(); snap.assign(w); snap.sync(); ts.append(time.perf_counter() - t0)
a = [int(x) for x in w.tick(5, hashes=True)[:, 0]]
b = [int(x) for x in snap.tick(5, hashes=True)[:, 0]]
print(json.dumps({"entities": w.count()[0], "dup_first_ms (allocates)": t_first * 1e3, "assign_ms_median": sorted(ts)[len(ts) // 2] * 1e3,
                  "copy_continues_identically": a == b}))
