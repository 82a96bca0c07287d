"""ctypes binding of libbittoss_b200.so (include/bittoss_b200.h) — the host-side mirror of
the reference's tick interface (ent.h:212-215) for Python callers.

There is no CPU path: importing works anywhere, but every compute call needs the CUDA
library built by `bittoss_b200.build()` and a visible GPU, and fails loudly otherwise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .desc import BtWorldDesc, WorldArrays

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_SO = os.environ.get("BT_LIB_SO", os.path.join(HERE, "libbittoss_b200.so"))  # override: A/B against another build
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

PHASE_CTRLS, PHASE_PREP, PHASE_GRAVITY, PHASE_PHYSICS, PHASE_FINISH, PHASE_ALL = 1, 2, 4, 8, 16, 31
PHASE_WORLD_RULES, PHASE_TICKS = 32, 64
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared"]

_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)


class BtError(RuntimeError):
    pass


class BtTickStats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in
                ("ticks", "iterations", "candidate_pairs", "positives", "pushes", "serial_pushes", "crushes",
                 "fumbles", "kernel_launches", "tree_builds", "tree_refits", "live_ents", "entity_ticks",
                 "pair_builds", "pairs_cached", "pair_overflows", "unsupported_events", "pair_patches", "order_ties")]


class BtEvent(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("kind", "handler", "world", "me_uid", "him_uid", "arg")]


class BtRequest(C.Structure):
    _fields_ = [("kind", C.c_int32), ("world", C.c_int32), ("me_uid", C.c_int32), ("other_uid", C.c_int32),
                ("a", C.c_int32 * 3)]


# request kinds (include/bittoss_b200.h BT_REQ_*)
REQ_DROP, REQ_DEAD, REQ_PICKUP, REQ_BUTTON, REQ_VEL, REQ_CENTER, REQ_SLIDER, REQ_ERASE, REQ_TYPEMASK, \
    REQ_COLLIDEMASK = range(10)


class BtProfile(C.Structure):
    _fields_ = [("n", C.c_int32), ("name", C.c_char_p * 24), ("launches", C.c_int64 * 24), ("ms", C.c_double * 24)]


def build(force=False, verbose=False):
    """Compile the CUDA library for sm_100a with nvcc (cross-compiles without a GPU).  Translation units compiled side by
    side: bt_world.cu (every kernel of the generic schedule + the host side of the C ABI) and bt_small.cu (the
    one-block-per-world step, by far the longest compile, in four variants); objects are kept in build/obj and reused
    while their sources are older."""
    hdrs = [os.path.join(CSRC, f) for f in ("bt_kernels.cuh", "bt_small.h", "bt_phys.cuh", "bt_raycast.cuh", "bt_types.h")]
    hdrs.append(os.path.join(INCLUDE, "bittoss_b200.h"))
    extra = os.environ.get("BT_NVCC_EXTRA", "").split()
    objdir = os.environ.get("BT_OBJ_DIR", os.path.join(os.path.dirname(CSRC), "..", "build", "obj"))
    os.makedirs(objdir, exist_ok=True)
    tag = ("_" + "_".join(extra).replace("/", "").replace("=", "")) if extra else ""
    # (source, object, extra flags): bt_small.cu once per (block size, resident warps per SM) variant of its kernel
    units = [(os.path.join(CSRC, n), os.path.join(objdir, n.rsplit(".", 1)[0] + tag + ".o"), []) for n in ("bt_world.cu", "bt_synth.cpp")]
    units += [(os.path.join(CSRC, "bt_small.cu"), os.path.join(objdir, "bt_small_%d_%d%s.o" % (bs, mw, tag)),
               ["-DSW_BS=%d" % bs, "-DSW_MIN_WARPS=%d" % mw]) for bs in (64, 128) for mw in (16, 32)]
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + extra + (["-Xptxas", "-v"] if verbose else [])
    procs = []
    for src, obj, flags in units:
        deps = [src] + hdrs
        if force or not os.path.exists(obj) or any(os.path.getmtime(obj) < os.path.getmtime(d) for d in deps):
            procs.append((src, subprocess.Popen(["nvcc"] + compile_flags + flags + ["-c", "-o", obj, src])))
    failed = [src for src, p in procs if p.wait() != 0]
    if failed:
        raise subprocess.CalledProcessError(1, "nvcc -c " + " ".join(failed))
    objs = [obj for _, obj, _ in units]
    if procs or not os.path.exists(LIB_SO) or any(os.path.getmtime(LIB_SO) < os.path.getmtime(o) for o in objs):
        subprocess.check_call(["nvcc"] + NVCC_FLAGS + ["-o", LIB_SO] + objs)
    return LIB_SO


_lib = None


def lib():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_SO):
        raise BtError("%s is missing: run `python -c 'import bittoss_b200; bittoss_b200.build()'` "
                      "(there is no CPU fallback)" % LIB_SO)
    L = C.CDLL(LIB_SO)
    L.bt_version.restype = C.c_char_p
    L.bt_last_error.restype = C.c_char_p
    L.bt_world_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    L.bt_world_destroy.argtypes = [C.c_void_p]
    L.bt_world_dup.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.bt_world_assign.argtypes = [C.c_void_p, C.c_void_p]
    L.bt_world_upload.argtypes = [C.c_void_p, C.POINTER(BtWorldDesc)]
    L.bt_world_count.argtypes = [C.c_void_p, _i32p, _i32p, _i32p]
    L.bt_world_download.argtypes = [C.c_void_p, C.POINTER(BtWorldDesc)]
    L.bt_world_step.argtypes = [C.c_void_p, C.c_uint32]
    L.bt_world_tick.argtypes = [C.c_void_p, C.c_int32, _u64p]
    L.bt_world_tick_async.argtypes = [C.c_void_p, C.c_int32]
    L.bt_world_sync.argtypes = [C.c_void_p]
    L.bt_world_tick_timed.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_float)]
    L.bt_world_apply_host_deltas.argtypes = [C.c_void_p, _i32p, _i32p, C.c_int32]
    L.bt_world_read_state.argtypes = [C.c_void_p, _i32p, _i32p, C.c_int32]
    L.bt_world_apply_host_deltas_async.argtypes = [C.c_void_p, _i32p, _i32p, C.c_int32]
    L.bt_world_send_host_deltas.argtypes = [C.c_void_p, _i32p, _i32p, C.c_int32]
    L.bt_world_commit_host_deltas.argtypes = [C.c_void_p]
    L.bt_world_read_state_async.argtypes = [C.c_void_p, _i32p, _i32p, C.c_int32, _i32p]
    L.bt_world_read_state_wait.argtypes = [C.c_void_p, C.c_int32]
    L.bt_world_stream.argtypes = [C.c_void_p]
    L.bt_world_stream.restype = C.c_void_p
    L.bt_world_hash.argtypes = [C.c_void_p, _u64p]
    if hasattr(L, "bt_world_hash_full"):
        L.bt_world_hash_full.argtypes = [C.c_void_p, _u64p]
    L.bt_world_stats.argtypes = [C.c_void_p, C.POINTER(BtTickStats), C.c_int]
    if hasattr(L, "bt_world_live_ents"):  # absent only from older builds named with BT_LIB_SO for an A/B run
        L.bt_world_live_ents.argtypes = [C.c_void_p]
        L.bt_world_live_ents.restype = C.c_int64
    L.bt_world_order_ties.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    L.bt_world_apply_requests.argtypes = [C.c_void_p, C.POINTER(BtRequest), C.c_int32]
    L.bt_world_drain_events.argtypes = [C.c_void_p, C.POINTER(BtEvent), C.c_int32, C.POINTER(C.c_int32)]
    L.bt_world_unsupported.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    L.bt_world_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.bt_world_profile_read.argtypes = [C.c_void_p, C.POINTER(BtProfile), C.c_int]
    L.bt_world_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.bt_raycast_batch.argtypes = [C.c_void_p, _i32p, _i32p, C.c_int32, _i32p]
    L.bt_raycast_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.bt_raycast_build.argtypes = [C.c_void_p]
    L.bt_synth_boxstack_size.argtypes = [C.c_int32, C.c_int32, _i32p]
    L.bt_synth_boxstack_fill.argtypes = [C.POINTER(BtWorldDesc), C.c_int32, C.c_int32, C.c_uint32, C.c_int32,
                                         C.c_int32]
    _lib = L
    return L


def _ck(rc, what):
    if rc != 0:
        raise BtError("%s failed with code %d: %s" % (what, rc, lib().bt_last_error().decode(errors="replace")))


def device_count():
    n = lib().bt_device_count()
    if n < 0:
        raise BtError("no usable CUDA device (code %d): %s" % (n, lib().bt_last_error().decode(errors="replace")))
    return n


def synth_boxstack(n_cubes, height=4, seed=777, push_every=64, push_max=60) -> WorldArrays:
    """BASELINE.json configs[2]: the synthetic box-stack world as plain host arrays."""
    n = C.c_int32()
    _ck(lib().bt_synth_boxstack_size(n_cubes, height, C.byref(n)), "bt_synth_boxstack_size")
    wa = WorldArrays(n.value, 0, 0, 1)
    _ck(lib().bt_synth_boxstack_fill(wa.ref(), n_cubes, height, seed, push_every, push_max), "bt_synth_boxstack_fill")
    return wa


def concat_worlds(worlds):
    """Concatenate single-world WorldArrays into one batched bt_world_desc (references rebased)."""
    n = sum(w.n for w in worlds)
    ns = sum(int(w.desc.n_sliders) for w in worlds)
    nw = sum(int(w.desc.n_wires) for w in worlds)
    out = WorldArrays(n, ns, nw, len(worlds))
    eb = sb = wb = 0
    ref_fields = ("holder", "holdee_head", "holdee_next", "new_holder")
    from .desc import _ENT_FIELDS, _SLIDER_FIELDS
    for g, w in enumerate(worlds):
        m, ms, mw = w.n, int(w.desc.n_sliders), int(w.desc.n_wires)
        for name, _, k in _ENT_FIELDS:
            src = w.a[name][: m * k]
            if name in ref_fields:
                src = np.where(src >= 0, src + eb, -1)
            out.a[name][eb * k: (eb + m) * k] = src
        out.a["world_id"][eb: eb + m] = g
        for name in _SLIDER_FIELDS:
            out.a[name][sb: sb + ms] = w.a[name][:ms]
        out.a["slider_off"][eb: eb + m + 1] = w.a["slider_off"][: m + 1] + sb
        out.a["wire_off"][eb: eb + m + 1] = w.a["wire_off"][: m + 1] + wb
        wt = w.a["wire_to"][:mw]
        out.a["wire_to"][wb: wb + mw] = np.where(wt >= 0, wt + eb, -1)
        for name in ("w_flip_flop", "w_gamerules", "w_rand"):
            out.a[name][g] = w.a[name][0]
        eb, sb, wb = eb + m, sb + ms, wb + mw
    return out


class World:
    """One device-resident gamestate batch. Mirrors the reference tick API:
    `step(phases)` ~ doUpdates(flushCtrls)/prepPhysics/doPhysics/finishStep (ent.h:212-215)."""

    def __init__(self, wa: WorldArrays = None, device=0, max_ents=None, max_sliders=None, max_wires=None,
                 max_worlds=None):
        L = lib()
        device_count()
        self.L = L
        self.h = C.c_void_p()
        me = max_ents if max_ents is not None else (wa.n if wa else 0)
        ms = max_sliders if max_sliders is not None else (int(wa.desc.n_sliders) if wa else 0)
        mw = max_wires if max_wires is not None else (int(wa.desc.n_wires) if wa else 0)
        mg = max_worlds if max_worlds is not None else (int(wa.desc.n_worlds) if wa else 1)
        _ck(L.bt_world_create(C.byref(self.h), device, me, ms, mw, mg), "bt_world_create")
        self.n_worlds = mg
        if wa is not None:
            self.upload(wa)

    def close(self):
        if getattr(self, "h", None):
            self.L.bt_world_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def dup(self):
        """dup(gs) (ent.cpp:1135-1200): an independent copy of this world, device to device."""
        other = World.__new__(World)
        other.L, other.h, other.n_worlds = self.L, C.c_void_p(), self.n_worlds
        _ck(self.L.bt_world_dup(self.h, C.byref(other.h)), "bt_world_dup")
        return other

    def assign(self, src):
        """self := src (same capacities), the rollback restore."""
        _ck(self.L.bt_world_assign(self.h, src.h), "bt_world_assign")
        self.n_worlds = src.n_worlds

    def upload(self, wa: WorldArrays):
        _ck(self.L.bt_world_upload(self.h, wa.ref()), "bt_world_upload")
        self.n_worlds = int(wa.desc.n_worlds)

    def step(self, phases=PHASE_ALL):
        _ck(self.L.bt_world_step(self.h, phases), "bt_world_step")

    def tick(self, n=1, hashes=False):
        if hashes:
            out = np.zeros(n * self.n_worlds, dtype=np.uint64)
            _ck(self.L.bt_world_tick(self.h, n, out.ctypes.data_as(_u64p)), "bt_world_tick")
            return out.reshape(n, self.n_worlds)
        _ck(self.L.bt_world_tick(self.h, n, None), "bt_world_tick")

    def tick_timed(self, n):
        """n whole steps; returns device milliseconds between CUDA events on the world's stream."""
        ms = C.c_float()
        _ck(self.L.bt_world_tick_timed(self.h, n, C.byref(ms)), "bt_world_tick_timed")
        return float(ms.value)

    def apply_host_deltas(self, d_center, d_vel):
        """uCenter / uVel accumulated by host tick handlers, (n,3) int32 arrays (may be pinned)."""
        n = d_center.size // 3
        _ck(self.L.bt_world_apply_host_deltas(self.h, C.cast(d_center.ctypes.data, _i32p),
                                              C.cast(d_vel.ctypes.data, _i32p), n), "bt_world_apply_host_deltas")

    def apply_requests(self, reqs):
        """deferred updater calls made by host-side handlers: [(kind, world, me_uid, other_uid, (a0, a1, a2)), ...]"""
        buf = (BtRequest * max(1, len(reqs)))()
        for i, (kind, world, me, other, a) in enumerate(reqs):
            buf[i].kind, buf[i].world, buf[i].me_uid, buf[i].other_uid = int(kind), int(world), int(me), int(other)
            for k in range(3):
                buf[i].a[k] = int(a[k]) if k < len(a) else 0
        _ck(self.L.bt_world_apply_requests(self.h, buf, len(reqs)), "bt_world_apply_requests")

    def read_state(self, center, vel):
        n = center.size // 3
        _ck(self.L.bt_world_read_state(self.h, C.cast(center.ctypes.data, _i32p), C.cast(vel.ctypes.data, _i32p), n),
            "bt_world_read_state")

    def apply_host_deltas_async(self, d_center, d_vel):
        """pipelined: the copy starts now on its own stream, the deltas are added in front of the next step (pinned arrays)"""
        n = d_center.size // 3
        _ck(self.L.bt_world_apply_host_deltas_async(self.h, C.cast(d_center.ctypes.data, _i32p),
                                                    C.cast(d_vel.ctypes.data, _i32p), n), "bt_world_apply_host_deltas_async")

    def send_host_deltas(self, d_center, d_vel):
        """pipelined, first half: only the copy starts (call it before the step the deltas must not reach)"""
        n = d_center.size // 3
        _ck(self.L.bt_world_send_host_deltas(self.h, C.cast(d_center.ctypes.data, _i32p), C.cast(d_vel.ctypes.data, _i32p), n),
            "bt_world_send_host_deltas")

    def commit_host_deltas(self):
        """pipelined, second half: the oldest batch sent is added in front of the next step"""
        _ck(self.L.bt_world_commit_host_deltas(self.h), "bt_world_commit_host_deltas")

    def read_state_async(self, center, vel):
        """pipelined: packs the state behind everything enqueued and starts the copy; returns the ticket for read_state_wait"""
        n = center.size // 3
        t = C.c_int32()
        _ck(self.L.bt_world_read_state_async(self.h, C.cast(center.ctypes.data, _i32p), C.cast(vel.ctypes.data, _i32p), n,
                                             C.byref(t)), "bt_world_read_state_async")
        return int(t.value)

    def read_state_wait(self, ticket):
        _ck(self.L.bt_world_read_state_wait(self.h, int(ticket)), "bt_world_read_state_wait")

    def sync(self):
        _ck(self.L.bt_world_sync(self.h), "bt_world_sync")

    def stream(self):
        return self.L.bt_world_stream(self.h)

    def hash(self):
        out = np.zeros(self.n_worlds, dtype=np.uint64)
        _ck(self.L.bt_world_hash(self.h, out.ctypes.data_as(_u64p)), "bt_world_hash")
        return [int(x) for x in out]

    def hash_full(self):
        """per-world hash of everything a .sav holds plus the flush state (bt_world_hash_full)"""
        out = np.zeros(self.n_worlds, dtype=np.uint64)
        _ck(self.L.bt_world_hash_full(self.h, out.ctypes.data_as(_u64p)), "bt_world_hash_full")
        return [int(x) for x in out]

    def count(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        _ck(self.L.bt_world_count(self.h, C.byref(a), C.byref(b), C.byref(c)), "bt_world_count")
        return a.value, b.value, c.value

    def download(self) -> WorldArrays:
        n, ns, nw = self.count()
        wa = WorldArrays(n, ns, nw, self.n_worlds)
        _ck(self.L.bt_world_download(self.h, wa.ref()), "bt_world_download")
        return wa

    def live_ents(self):
        """live entities after the last finished step (host-side bookkeeping, no synchronisation)"""
        return int(self.L.bt_world_live_ents(self.h))

    def stats(self, reset=False):
        s = BtTickStats()
        _ck(self.L.bt_world_stats(self.h, C.byref(s), int(reset)), "bt_world_stats")
        return {k: int(getattr(s, k)) for k, _ in BtTickStats._fields_}

    def drain_events(self, max_events=4096):
        """host-side handler events (PUSH_TEAMSELECT / PUSH_LOADBLOCK, crush handlers) since the last drain"""
        buf = (BtEvent * max_events)()
        n = C.c_int32()
        _ck(self.L.bt_world_drain_events(self.h, buf, max_events, C.byref(n)), "bt_world_drain_events")
        return [{k: int(getattr(buf[i], k)) for k, _ in BtEvent._fields_} for i in range(n.value)]

    def order_ties_per_world(self):
        out = np.zeros(self.n_worlds, dtype=np.uint32)
        _ck(self.L.bt_world_order_ties(self.h, out.ctypes.data_as(C.POINTER(C.c_uint32))), "bt_world_order_ties")
        return out

    def unsupported_per_world(self):
        out = np.zeros(self.n_worlds, dtype=np.uint32)
        _ck(self.L.bt_world_unsupported(self.h, out.ctypes.data_as(C.POINTER(C.c_uint32))), "bt_world_unsupported")
        return out

    def profile_enable(self, on=True):
        _ck(self.L.bt_world_profile_enable(self.h, int(on)), "bt_world_profile_enable")

    def profile(self, reset=False):
        p = BtProfile()
        _ck(self.L.bt_world_profile_read(self.h, C.byref(p), int(reset)), "bt_world_profile_read")
        return {p.name[i].decode(): (int(p.launches[i]), float(p.ms[i])) for i in range(p.n)}

    def set_option(self, name, value):
        _ck(self.L.bt_world_set_option(self.h, name.encode(), int(value)), "bt_world_set_option")

    def raycast(self, origin, look):
        """selectCast (raycast.h:1) batched: origin = casting entity index per ray, look = (n,3)."""
        origin = np.ascontiguousarray(origin, dtype=np.int32)
        look = np.ascontiguousarray(look, dtype=np.int32)
        out = np.empty(len(origin), dtype=np.int32)
        _ck(self.L.bt_raycast_batch(self.h, origin.ctypes.data_as(_i32p), look.ctypes.data_as(_i32p), len(origin),
                                    out.ctypes.data_as(_i32p)), "bt_raycast_batch")
        return out

    def raycast_device(self, d_origin_ptr, d_look_ptr, n_rays, d_winner_ptr):
        _ck(self.L.bt_raycast_batch_device(self.h, d_origin_ptr, d_look_ptr, n_rays, d_winner_ptr),
            "bt_raycast_batch_device")

    def raycast_build(self):
        _ck(self.L.bt_raycast_build(self.h), "bt_raycast_build")
