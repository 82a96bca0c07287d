"""ctypes binding of tests/emu/libbtemu.so — the host single-stepper of the kernel bodies.

TEST TOOLING ONLY (see tests/emu/bt_emu.cpp).  Never imported by the product or bench.py.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from bittoss_b200.desc import BtWorldDesc, WorldArrays  # noqa: E402

EMU_SO = os.environ.get("BT_EMU_SO", os.path.join(HERE, "emu", "libbtemu.so"))  # override: a sanitizer build
_i32p = C.POINTER(C.c_int32)


def build_emu(force=False):
    src = os.path.join(HERE, "emu", "bt_emu.cpp")
    deps = [src] + [os.path.join(ROOT, "bittoss_b200", "csrc", f) for f in ("bt_phys.cuh", "bt_types.h", "bt_raycast.cuh")]
    if "BT_EMU_SO" in os.environ:
        return
    if not force and os.path.exists(EMU_SO) and all(os.path.getmtime(EMU_SO) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", EMU_SO, src])


class EmuWorld:
    def __init__(self, wa: WorldArrays, rebuild_every=4, shuffle_seed=0, lenient=False):
        build_emu()
        L = C.CDLL(EMU_SO)
        L.bte_create.restype = C.c_void_p
        L.bte_create.argtypes = [C.POINTER(BtWorldDesc), C.c_int, C.c_int]
        L.bte_destroy.argtypes = [C.c_void_p]
        L.bte_step.argtypes = [C.c_void_p, C.c_uint32]
        L.bte_hash.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.bte_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.bte_peek.argtypes = [C.c_void_p, _i32p, _i32p, C.POINTER(C.c_uint8)]
        L.bte_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        L.bte_pair_rebuilds.argtypes = [C.c_void_p]
        L.bte_pair_rebuilds.restype = C.c_int64
        L.bte_pair_patches.argtypes = [C.c_void_p]
        L.bte_pair_patches.restype = C.c_int64
        L.bte_events.argtypes = [C.c_void_p]
        L.bte_events.restype = C.c_int64
        L.bte_unsupported.argtypes = [C.c_void_p]
        L.bte_unsupported.restype = C.c_int64
        L.bte_order_ties.argtypes = [C.c_void_p]
        L.bte_order_ties.restype = C.c_int64
        L.bte_raycast.argtypes = [C.c_void_p, _i32p, _i32p, C.c_int32, _i32p]
        L.bte_drain_events.argtypes = [C.c_void_p, _i32p, C.c_int]
        L.bte_request.argtypes = [C.c_void_p, C.c_int] + [C.c_int32] * 5
        L.bte_export_misc.argtypes = [C.c_void_p] + [_i32p] * 6 + [C.POINTER(C.c_uint32), C.c_int, C.c_int]
        self.L = L
        self.n_worlds = int(wa.desc.n_worlds)
        self.n = wa.n
        self.h = L.bte_create(wa.ref(), rebuild_every, shuffle_seed)
        if lenient:
            self.set_option("allow_unsupported", 1)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.bte_destroy(self.h)
            self.h = None

    def step(self, phases=31):
        rc = self.L.bte_step(self.h, phases)
        if rc != 0:
            raise RuntimeError("bte_step rc=%d" % rc)

    def set_option(self, name, value):
        self.L.bte_set_option(self.h, name.encode(), int(value))

    def events(self):
        return int(self.L.bte_events(self.h))

    def drain_events(self):
        buf = (C.c_int32 * (5 * 4096))()
        n = self.L.bte_drain_events(self.h, buf, 4096)
        return [dict(kind=buf[5 * i], handler=buf[5 * i + 1], me_uid=buf[5 * i + 2], him_uid=buf[5 * i + 3],
                     arg=buf[5 * i + 4]) for i in range(min(n, 4096))]

    def request(self, kind, me_uid, other_uid=-1, a=(0, 0, 0)):
        a = list(a) + [0] * (3 - len(a))
        rc = self.L.bte_request(self.h, kind, me_uid, other_uid, a[0], a[1], a[2])
        if rc != 0:
            raise RuntimeError("bte_request rc=%d" % rc)

    def export_misc(self):
        """-> {uid: (holder uid, [sliders], [wire target uids])} of live entities"""
        me, mi = self.n + 1, 1 << 16
        uid, hol, ns, nw = [(C.c_int32 * me)() for _ in range(4)]
        sl, wr = (C.c_int32 * mi)(), (C.c_int32 * mi)()
        bt = (C.c_uint32 * (2 * me))()
        k = self.L.bte_export_misc(self.h, uid, hol, ns, sl, nw, wr, bt, me, mi)
        assert k >= 0
        out, so, wo = {}, 0, 0
        for i in range(k):
            btn = [(int(bt[2 * i + q]) >> (8 * b)) & 0xff for q in range(2) for b in range(4)]
            out[int(uid[i])] = (int(hol[i]), [int(sl[so + j]) for j in range(ns[i])], [int(wr[wo + j]) for j in range(nw[i])], btn)
            so += ns[i]
            wo += nw[i]
        return out

    def unsupported(self):
        return int(self.L.bte_unsupported(self.h))

    def order_ties(self):
        return int(self.L.bte_order_ties(self.h))

    def pair_rebuilds(self):
        return int(self.L.bte_pair_rebuilds(self.h))

    def pair_patches(self):
        return int(self.L.bte_pair_patches(self.h))

    def hash(self):
        out = (C.c_uint64 * self.n_worlds)()
        self.L.bte_hash(self.h, out)
        return [int(x) for x in out]

    def stats(self):
        out = (C.c_int64 * 8)()
        self.L.bte_stats(self.h, out)
        keys = ["ticks", "iterations", "candidate_pairs", "positives", "pushes", "serial_pushes", "crushes", "fumbles"]
        return dict(zip(keys, [int(x) for x in out]))

    def peek(self):
        c = np.zeros(3 * self.n, dtype=np.int32)
        v = np.zeros(3 * self.n, dtype=np.int32)
        a = np.zeros(self.n, dtype=np.uint8)
        self.L.bte_peek(self.h, c.ctypes.data_as(_i32p), v.ctypes.data_as(_i32p), a.ctypes.data_as(C.POINTER(C.c_uint8)))
        return c.reshape(-1, 3), v.reshape(-1, 3), a

    def raycast(self, origin, look):
        origin = np.ascontiguousarray(origin, dtype=np.int32)
        look = np.ascontiguousarray(look, dtype=np.int32)
        out = np.empty(len(origin), dtype=np.int32)
        rc = self.L.bte_raycast(self.h, origin.ctypes.data_as(_i32p), look.ctypes.data_as(_i32p), len(origin),
                                out.ctypes.data_as(_i32p))
        if rc != 0:
            raise RuntimeError("bte_raycast rc=%d" % rc)
        return out
