"""Reference hash stream of BASELINE.json configs[2] AT ITS NAMED SIZE: the 1 000 000-cube box-stack world
(seed 777, height 4, a push on every 64th column, |v| <= 60: exactly what bench.py runs on rank 0) stepped by
the UNMODIFIED reference (oracle/_ref, built from /root/reference by oracle/Makefile), one state hash per tick.

    python tests/golden/make_stack1m_ref.py [ticks=160] [cubes=1000000] [out] [push_every=64]

push_every=1 (tests/golden/stack1m_pushall_ref_hashes.npz) kicks the top cube of EVERY column: the "dynamic"
phase bench.py reports, in which pair lists are rebuilt or patched while the kicked cubes slide.

About 5 s per tick on one core of the build container.  The fixture pins the CUDA path at full size:
tests/test_gpu_parity.py::test_cuda_1m_stack_vs_reference and bench.py's `parity` record compare their own
per-tick hashes with this stream.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pybind  # noqa: E402


def main():
    ticks = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    cubes = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "tests", "golden", "stack1m_ref_hashes.npz")
    push_every = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    R = pybind.RefLib()
    R.new_world()
    with pybind.stdout_to_stderr():
        n = R.make_boxstack(cubes, 4, 777, push_every, 60)
    hashes, live, secs = [], [], []
    for t in range(ticks):
        t0 = time.time()
        with pybind.stdout_to_stderr():
            R.step(1)
        hashes.append(R.hash())
        live.append(R.count()[0])
        secs.append(time.time() - t0)
        if t % 10 == 0:
            print("tick %d  %.2f s  live %d  hash %016x" % (t, secs[-1], live[-1], hashes[-1]), flush=True)
            np.savez_compressed(out, hashes=np.array(hashes, dtype=np.uint64), live=np.array(live, dtype=np.int64),
                                cubes=cubes, height=4, seed=777, push_every=push_every, push_max=60, entities=n,
                                seconds=np.array(secs))
    np.savez_compressed(out, hashes=np.array(hashes, dtype=np.uint64), live=np.array(live, dtype=np.int64),
                        cubes=cubes, height=4, seed=777, push_every=push_every, push_max=60, entities=n,
                        seconds=np.array(secs))
    print("wrote", out)


if __name__ == "__main__":
    main()
