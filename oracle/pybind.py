"""ctypes bindings of the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* `RefLib`    — oracle/_ref/libbtref.so : the unmodified reference sim core behind
                oracle/ref_harness.cpp (built from /root/reference when present).
* `OracleLib` — oracle/libbtoracle.so   : the plain-C restatement (oracle/bt_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from bittoss_b200.desc import BtWorldDesc, WorldArrays  # noqa: E402  (plain buffer plumbing)

REF_SO = os.environ.get("BTREF_SO", os.path.join(HERE, "_ref", "libbtref.so"))
REF_DATA_PARENT = os.path.join(HERE, "_ref")
ORACLE_SO = os.path.join(HERE, "libbtoracle.so")

_i32p = C.POINTER(C.c_int32)


def build(ref: bool = True):
    """Compile the checkers (the reference only if its tree is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "libbtoracle.so"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


import contextlib


@contextlib.contextmanager
def stdout_to_stderr():
    """The reference prints diagnostics with puts()/printf() (file.cpp:16, ent.cpp:519,582);
    keep them away from a caller's machine-readable stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        yield
    finally:
        for name in (None, "libc.so.6"):
            try:
                C.CDLL(name).fflush(None)
                break
            except Exception:
                continue
        os.dup2(saved, 1)
        os.close(saved)


def have_ref():
    return os.path.exists(REF_SO) and os.path.isdir(os.path.join(REF_DATA_PARENT, "data"))


class RefLib:
    """The reference itself. One per process (the reference keeps file-static state)."""

    _inst = None

    def __new__(cls):
        if cls._inst is None:
            cls._inst = super().__new__(cls)
            cls._inst._load()
        return cls._inst

    def _load(self):
        if not have_ref():
            raise RuntimeError("oracle/_ref is not built (run `make -C oracle`)")
        L = C.CDLL(REF_SO)
        L.btref_init.argtypes = [C.c_char_p]
        L.btref_load_sav.argtypes = [C.c_char_p, C.c_int, _i32p, C.c_int32]
        L.btref_set_options.argtypes = [C.c_int, C.c_uint32]
        L.btref_serialize.argtypes = [C.c_char_p, C.c_int]
        L.btref_make_boxstack.argtypes = [C.c_int32, C.c_int32, C.c_uint32, C.c_int32, C.c_int32]
        L.btref_strip_handlers.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.btref_kick.argtypes = [C.c_uint32, C.c_int32, C.c_int32]
        L.btref_count.argtypes = [_i32p, _i32p, _i32p]
        L.btref_export.argtypes = [C.POINTER(BtWorldDesc)]
        L.btref_build_from_desc.argtypes = [C.POINTER(BtWorldDesc)]
        L.btref_export_at.argtypes = [C.POINTER(BtWorldDesc), C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.btref_hash.restype = C.c_uint64
        L.btref_hash_full.restype = C.c_uint64
        L.btref_selectcast_batch.argtypes = [_i32p, _i32p, C.c_int32, _i32p]
        L.btref_time_steps.argtypes = [C.c_int, C.c_int32, C.POINTER(C.c_int64)]
        L.btref_time_steps.restype = C.c_double
        L.btref_add_ent.argtypes = [_i32p, _i32p, _i32p, C.c_uint32, C.c_uint32, C.c_int]
        L.btref_lab_set_handlers.argtypes = [C.c_int32, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, C.c_int]
        L.btref_lab_set_sliders.argtypes = [C.c_int32, C.c_int, _i32p]
        L.btref_lab_add_wire.argtypes = [C.c_int32, C.c_int32]
        L.btref_lab_set_button.argtypes = [C.c_int32, C.c_int, C.c_int]
        L.btref_lab_request.argtypes = [C.c_int32, C.c_int, C.c_int32, C.c_int32]
        L.btref_lab_crush_log.argtypes = [_i32p, C.c_int]
        L.btref_lab_request3.argtypes = [C.c_int32, C.c_int, C.c_int32, C.c_int32, C.c_int32]
        with stdout_to_stderr():
            rc = L.btref_init(REF_DATA_PARENT.encode())
        if rc != 0:
            raise RuntimeError("btref_init failed")
        self.L = L

    # -- worlds --
    def new_world(self):
        return self.L.btref_new_world()

    def select(self, wid):
        assert self.L.btref_select(wid) == 0

    def free_world(self, wid):
        self.L.btref_free_world(wid)

    def load_sav(self, name, players=(), rand_seed=12345):
        arr = (C.c_int32 * max(1, len(players)))(*players)
        if self.L.btref_load_sav(name.encode(), len(players), arr, rand_seed) != 0:
            raise RuntimeError("cannot load %s" % name)

    def build_from(self, wa):
        """the current (empty) world := the entities of a WorldArrays, made by the reference's own initEnt"""
        return int(self.L.btref_build_from_desc(wa.ref()))

    def make_boxstack(self, n_cubes, height=4, seed=777, push_every=64, push_max=60):
        return self.L.btref_make_boxstack(n_cubes, height, seed, push_every, push_max)

    def strip_handlers(self, drop_ticks=True, keep_push=0b101, keep_pushed=0b1101, keep_pair=0b1001,
                       keep_crush=0b1):
        self.L.btref_strip_handlers(int(drop_ticks), keep_push, keep_pushed, keep_pair, keep_crush)

    def kick(self, seed, max_v, every=1):
        self.L.btref_kick(seed, max_v, every)

    # -- handler lab (small hand-built worlds, see tests/golden/make_handler_labs.py) --
    def add_ent(self, center, vel, radius, type_mask, collide_mask, who_moves):
        a = lambda v: (C.c_int32 * 3)(*[int(x) for x in v])
        return int(self.L.btref_add_ent(a(center), a(vel), a(radius), type_mask, collide_mask, who_moves))

    def lab_set_handlers(self, uid, push=-1, pushed=-1, crush=-1, pair=(-1, -1, -1, -1), tick=-1, tick_held=-1):
        p = (C.c_int32 * 4)(*pair)
        assert self.L.btref_lab_set_handlers(uid, push, pushed, crush, p, tick, tick_held) == 0

    def lab_set_sliders(self, uid, values):
        v = (C.c_int32 * max(1, len(values)))(*[int(x) for x in values])
        assert self.L.btref_lab_set_sliders(uid, len(values), v) == 0

    def lab_add_wire(self, a, b):
        assert self.L.btref_lab_add_wire(a, b) == 0

    def lab_set_button(self, uid, ix, v=1):
        assert self.L.btref_lab_set_button(uid, ix, v) == 0

    def lab_request(self, uid, what, other=-1, arg=0):
        """what: 0 uDrop, 1 uDead, 2 uPickup(uid picks up other, arg = hold flags)"""
        assert self.L.btref_lab_request(uid, what, other, arg) == 0

    def lab_real_crush(self, on=True):
        self.L.btref_lab_real_crush(int(on))

    def lab_record_all_crushes(self):
        return int(self.L.btref_lab_record_all_crushes())

    def lab_request3(self, uid, kind, a=(0, 0, 0)):
        """kind = BT_REQ_VEL / CENTER / SLIDER / ERASE / TYPEMASK / COLLIDEMASK (include/bittoss_b200.h)"""
        a = list(a) + [0] * (3 - len(a))
        assert self.L.btref_lab_request3(uid, kind, int(a[0]), int(a[1]), int(a[2])) == 0

    def lab_crush_log(self):
        buf = (C.c_int32 * 4096)()
        n = self.L.btref_lab_crush_log(buf, 4096)
        return [int(buf[i]) for i in range(min(n, 4096))]

    def set_options(self, spacetrash_on=0, input_key=0):
        self.L.btref_set_options(spacetrash_on, input_key)

    def set_gamerules(self, r):
        self.L.btref_set_gamerules(r)

    def get_gamerules(self):
        return self.L.btref_get_gamerules()

    def serialize(self):
        """bytes the reference's own `serialize` writes for the current world"""
        cap = 1 << 20
        while True:
            buf = C.create_string_buffer(cap)
            n = self.L.btref_serialize(buf, cap)
            if n <= cap:
                return buf.raw[:n]
            cap = n

    # -- stepping --
    def step(self, mode=1):
        self.L.btref_step(mode)

    def pre_physics(self, mode=1):
        self.L.btref_pre_physics(mode)

    def do_physics(self):
        self.L.btref_do_physics()

    def hash(self):
        return int(self.L.btref_hash())

    def hash_full(self):
        """bt_world_hash_full restated over the reference's `ent`s (everything a .sav holds + the flush state)"""
        return int(self.L.btref_hash_full())

    def count(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self.L.btref_count(C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def export(self) -> WorldArrays:
        n, ns, nw = self.count()
        wa = WorldArrays(n, ns, nw, 1)
        assert self.L.btref_export(wa.ref()) == 0
        return wa

    def selectcast(self, origin, look):
        origin = np.ascontiguousarray(origin, dtype=np.int32)
        look = np.ascontiguousarray(look, dtype=np.int32)
        out = np.empty(len(origin), dtype=np.int32)
        self.L.btref_selectcast_batch(origin.ctypes.data_as(_i32p), look.ctypes.data_as(_i32p), len(origin),
                                      out.ctypes.data_as(_i32p))
        return out

    def time_steps(self, n, mode=1):
        et = C.c_int64()
        s = self.L.btref_time_steps(mode, n, C.byref(et))
        return s, et.value


class OracleLib:
    """Plain-C restatement of the path (oracle/bt_oracle.c)."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.bto_create.restype = C.c_void_p
        L.bto_create.argtypes = [C.POINTER(BtWorldDesc)]
        L.bto_destroy.argtypes = [C.c_void_p]
        L.bto_step.argtypes = [C.c_void_p, C.c_uint32]
        L.bto_hash.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.bto_count.argtypes = [C.c_void_p, _i32p, _i32p, _i32p]
        L.bto_export.argtypes = [C.c_void_p, C.POINTER(BtWorldDesc)]
        L.bto_selectcast_batch.argtypes = [C.c_void_p, _i32p, _i32p, C.c_int32, _i32p]
        L.bto_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.bto_time_steps.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]
        L.bto_time_steps.restype = C.c_double
        L.bto_crush_log.argtypes = [C.c_void_p, _i32p, C.c_int]
        L.bto_order_ties.argtypes = [C.c_void_p]
        L.bto_order_ties.restype = C.c_int64
        L.bto_request.argtypes = [C.c_void_p, C.c_int] + [C.c_int32] * 5
        self.L = L

    def create(self, wa: WorldArrays):
        h = self.L.bto_create(wa.ref())
        if not h:
            raise RuntimeError("bto_create failed")
        return OracleWorld(self.L, h, int(wa.desc.n_worlds))


class OracleWorld:
    def __init__(self, L, h, n_worlds):
        self.L, self.h, self.n_worlds = L, h, n_worlds

    def __del__(self):
        if getattr(self, "h", None):
            self.L.bto_destroy(self.h)
            self.h = None

    def step(self, phases=31):
        rc = self.L.bto_step(self.h, phases)
        if rc != 0:
            raise RuntimeError("bto_step rc=%d" % rc)

    def hash(self):
        out = (C.c_uint64 * self.n_worlds)()
        self.L.bto_hash(self.h, out)
        return [int(x) for x in out]

    def request(self, kind, me_uid, other_uid=-1, a=(0, 0, 0)):
        a = list(a) + [0] * (3 - len(a))
        if self.L.bto_request(self.h, kind, me_uid, other_uid, a[0], a[1], a[2]) != 0:
            raise RuntimeError("bto_request failed")

    def order_ties(self):
        return int(self.L.bto_order_ties(self.h))

    def crush_log(self):
        buf = (C.c_int32 * 4096)()
        n = self.L.bto_crush_log(self.h, buf, 4096)
        return [int(buf[i]) for i in range(min(n, 4096))]

    def count(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self.L.bto_count(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def export(self) -> WorldArrays:
        n, ns, nw = self.count()
        wa = WorldArrays(n, ns, nw, self.n_worlds)
        assert self.L.bto_export(self.h, wa.ref()) == 0
        return wa

    def selectcast(self, origin, look):
        origin = np.ascontiguousarray(origin, dtype=np.int32)
        look = np.ascontiguousarray(look, dtype=np.int32)
        out = np.empty(len(origin), dtype=np.int32)
        self.L.bto_selectcast_batch(self.h, origin.ctypes.data_as(_i32p), look.ctypes.data_as(_i32p),
                                    len(origin), out.ctypes.data_as(_i32p))
        return out

    def stats(self):
        out = (C.c_int64 * 8)()
        self.L.bto_stats(self.h, out)
        keys = ["ticks", "iterations", "candidate_pairs", "positives", "pushes", "serial_pushes", "crushes",
                "fumbles"]
        return dict(zip(keys, [int(x) for x in out]))

    def time_steps(self, n):
        et = C.c_int64()
        s = self.L.bto_time_steps(self.h, n, C.byref(et))
        return s, et.value
