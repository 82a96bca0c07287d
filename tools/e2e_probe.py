"""Where the end-to-end step time goes: the pipelined exchange with parts switched off (1 M-cube world, 30 steps each)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bittoss_b200
w = bittoss_b200.World(bittoss_b200.synth_boxstack(1000000, 4, 777, 64, 60))
w.tick(55)
live = w.count()[0]
bufs = [[torch.zeros((live, 3), dtype=torch.int32).pin_memory().numpy() for _ in range(4)] for _ in range(2)]
def run(h2d, d2h, hash_, n=30, hash_first=False):
    tickets = [None, None]
    if h2d: w.apply_host_deltas_async(bufs[0][0], bufs[0][1])
    w.sync()
    t0 = time.perf_counter()
    for t in range(n):
        k = t & 1
        if h2d and t + 1 < n: w.send_host_deltas(bufs[k ^ 1][0], bufs[k ^ 1][1])
        w.step(31)
        if hash_ and hash_first: w.hash()
        if d2h: tickets[k] = w.read_state_async(bufs[k][2], bufs[k][3])
        if h2d and t + 1 < n: w.commit_host_deltas()
        if hash_ and not hash_first: w.hash()
        if d2h and tickets[k ^ 1] is not None:
            w.read_state_wait(tickets[k ^ 1]); tickets[k ^ 1] = None
    w.sync()
    return 1e3 * (time.perf_counter() - t0) / n
for name, args in (("tick only", (0, 0, 0)), ("tick + hash", (0, 0, 1)), ("h2d", (1, 0, 0)), ("h2d + hash", (1, 0, 1)), ("d2h", (0, 1, 0)),
                   ("d2h + hash", (0, 1, 1)), ("h2d + d2h", (1, 1, 0)), ("all", (1, 1, 1))):
    run(*args, n=4)
    print("%-12s %.3f ms/step   (hash first: %.3f)" % (name, run(*args), run(*args, hash_first=True)), flush=True)
def serial(n=30):
    t0 = time.perf_counter()
    for _ in range(n):
        w.apply_host_deltas(bufs[0][0], bufs[0][1]); w.step(31); w.read_state(bufs[0][2], bufs[0][3]); w.hash()
    return 1e3 * (time.perf_counter() - t0) / n
serial(3); print("serial       %.3f ms/step" % serial())
w.close()
