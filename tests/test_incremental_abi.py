"""The incremental part of the C ABI (include/bittoss_b200.h: bt_world_grow / poke / pull / extent / set_wires /
next_root_phase / hash_full), called directly through ctypes — on the kernel-body emulator (`-m "not gpu"`, it exports
the same entry points) and on libbittoss_b200.so (`-m gpu`).  The host binding (host/bt_glue.cpp) is built on exactly
these calls; here they are pinned one by one against the oracle and the reference."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bittoss_b200.desc import BtWorldDesc, WorldArrays  # noqa: E402
from util import load_scenario  # noqa: E402

# BT_A_* / BT_OP_* of include/bittoss_b200.h
(A_CENTER, A_VEL, A_OLD, A_DCENTER, A_DVEL, A_RADIUS, A_TYPE, A_COLLIDE, A_NTYPE, A_NCOLLIDE, A_FRICTION, A_HOLDER,
 A_HOLDFLAGS, A_HOLDROOT, A_HOLDEE, A_LLN, A_LLP, A_NEWHOLDER, A_NEWHOLDFLAGS, A_UID, A_WORLD, A_HPHYS, A_HPAIR, A_HTICK,
 A_BTN_V, A_BTN_V2, A_NEWDROP, A_DEADMAX0, A_DEADMAX1, A_FULLFLUSH, A_DEAD, A_INLIST, A_ROOTKEY, A_SLIDER_OFF, A_WIRE_OFF,
 A_WIRE_CAP, A_WIRE_CNT, A_SV, A_SMIN, A_SMAX, A_WIRE_TO, A_W_FLIP, A_W_GAMERULES, A_W_RAND, A_DEADLIST, A_NUM) = range(46)
OP_SPAWN, OP_UWIRE, OP_UUNWIRE = 100, 101, 102


class Poke(C.Structure):
    _fields_ = [("array", C.c_int32), ("index", C.c_int32), ("v", C.c_int32 * 3)]


class Abi:
    """the entry points, from whichever library serves them"""

    def __init__(self, so):
        L = C.CDLL(so)
        vp, i32, i32p = C.c_void_p, C.c_int32, C.POINTER(C.c_int32)
        L.bt_last_error.restype = C.c_char_p
        L.bt_world_create.argtypes = [C.POINTER(vp), C.c_int, i32, i32, i32, i32]
        L.bt_world_destroy.argtypes = [vp]
        L.bt_world_upload.argtypes = [vp, C.POINTER(BtWorldDesc)]
        L.bt_world_step.argtypes = [vp, C.c_uint32]
        L.bt_world_hash.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.bt_world_hash_full.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.bt_world_poke.argtypes = [vp, C.POINTER(Poke), i32]
        L.bt_world_pull.argtypes = [vp, i32, i32, i32, vp]
        L.bt_world_extent.argtypes = [vp, i32, i32p]
        L.bt_world_grow.argtypes = [vp, i32, i32p, i32p, i32, i32p]
        L.bt_world_set_wires.argtypes = [vp, i32, i32p, i32, i32]
        L.bt_world_next_root_phase.argtypes = [vp]
        L.bt_world_next_root_phase.restype = C.c_int64
        self.L = L

    def world(self, wa, cap_ents, cap_sliders=64, cap_wires=64):
        h = C.c_void_p()
        assert self.L.bt_world_create(C.byref(h), 0, cap_ents, cap_sliders, cap_wires, 1) == 0, self.L.bt_last_error()
        assert self.L.bt_world_upload(h, wa.ref()) == 0, self.L.bt_last_error()
        return h

    def ck(self, rc):
        assert rc == 0, (rc, self.L.bt_last_error())

    def hash(self, h, full=False):
        out = (C.c_uint64 * 1)()
        self.ck((self.L.bt_world_hash_full if full else self.L.bt_world_hash)(h, out))
        return int(out[0])

    def extent(self, h, a):
        n = C.c_int32()
        self.ck(self.L.bt_world_extent(h, a, C.byref(n)))
        return n.value

    def pull(self, h, a, dtype, per=1):
        n = self.extent(h, a)
        out = np.zeros(n * per, dtype=dtype)
        self.ck(self.L.bt_world_pull(h, a, 0, n, out.ctypes.data_as(C.c_void_p)))
        return out.reshape(n, per) if per > 1 else out

    def poke(self, h, rows):
        buf = (Poke * len(rows))()
        for k, (a, i, v) in enumerate(rows):
            v = list(v) + [0] * (3 - len(v))
            buf[k].array, buf[k].index = a, i
            for q in range(3):
                buf[k].v[q] = int(np.int32(np.uint32(v[q] & 0xffffffff)))
        self.ck(self.L.bt_world_poke(h, buf, len(rows)))


def _so(backend):
    if backend == "emu":
        import emubind
        emubind.build_emu()
        return emubind.EMU_SO
    import bittoss_b200
    bittoss_b200.device_count()
    return bittoss_b200.capi.LIB_SO


def _with_extra_cubes(wa, cubes):
    """the same world with `cubes` = [(center, vel)] appended as the newest entities (what a host's initEnt makes)"""
    n = wa.n
    out = WorldArrays(n + len(cubes), int(wa.desc.n_sliders), int(wa.desc.n_wires), 1)
    for k, arr in wa.a.items():
        out.a[k][: len(arr)] = arr
    for j, (c, v) in enumerate(cubes):
        i = n + j
        out.a["uid"][i] = i
        out.a["center"][3 * i: 3 * i + 3] = c
        out.a["vel"][3 * i: 3 * i + 3] = v
        out.a["old"][3 * i: 3 * i + 3] = np.array(c) - np.array(v)
        out.a["radius"][3 * i: 3 * i + 3] = 400
        out.a["type_mask"][i] = out.a["new_type_mask"][i] = 5
        out.a["collide_mask"][i] = out.a["new_collide_mask"][i] = 6
        out.a["friction"][i] = 16
        for f in ("holder", "holdee_head", "holdee_next", "new_holder"):
            out.a[f][i] = -1
        out.a["root_order"][i] = 0
        out.a["handlers"][10 * i] = 6          # WHOMOVES_WOOD
        out.a["slider_off"][i + 1] = out.a["slider_off"][i]
        out.a["wire_off"][i + 1] = out.a["wire_off"][i]
    roots = out.a["root_order"][:n]
    roots[roots >= 0] += len(cubes)            # the newcomers are the head of gs->rootEnts (ent.cpp:60-66)
    for j in range(len(cubes)):
        out.a["root_order"][n + j] = len(cubes) - 1 - j
    return out


def _grow_poke_spawn(backend, oracle_lib):
    wa, _ = load_scenario("stack512")
    cubes = [((300, 200, -6000), (0, 0, 90)), ((5300, -150, -4500), (12, -7, 40))]
    want = oracle_lib.create(_with_extra_cubes(wa, cubes))
    abi = Abi(_so(backend))
    h = abi.world(wa, wa.n + 16)
    first = C.c_int32()
    ns, wc = (C.c_int32 * 2)(0, 0), (C.c_int32 * 2)(2, 2)
    abi.ck(abi.L.bt_world_grow(h, 2, ns, wc, 0, C.byref(first)))
    assert first.value == wa.n and abi.extent(h, A_CENTER) == wa.n + 2
    assert not abi.pull(h, A_INLIST, np.uint8)[wa.n:].any()          # out of the world until BT_OP_SPAWN
    phase = abi.L.bt_world_next_root_phase(h)
    rows = []
    for j, (c, v) in enumerate(cubes):
        s = wa.n + j
        key = (phase << 37) | (j + 1)                                  # later creation = nearer the head
        rows += [(A_CENTER, s, c), (A_VEL, s, v), (A_OLD, s, tuple(np.array(c) - np.array(v))), (A_RADIUS, s, (400, 400, 400)),
                 (A_TYPE, s, (5,)), (A_COLLIDE, s, (6,)), (A_NTYPE, s, (5,)), (A_NCOLLIDE, s, (6,)), (A_FRICTION, s, (16,)),
                 (A_UID, s, (s,)), (A_HPHYS, s, (6,)), (A_ROOTKEY, s, (key & 0xffffffff, key >> 32)), (OP_SPAWN, s, ())]
    abi.poke(h, rows)
    assert abi.hash(h) == want.hash()[0]
    for t in range(25):
        want.step(31)
        abi.ck(abi.L.bt_world_step(h, 31))
        assert abi.hash(h) == want.hash()[0], "tick %d" % t
        if t == 9:   # uVel on the first newcomer: a d_vel write here, the oracle's own updater there
            want.request(4, wa.n, -1, (25, -10, -60))
            abi.poke(h, [(A_DVEL, wa.n, (25, -10, -60))])
    # what comes back: centres of every live slot equal the oracle's export (no deaths in this scenario)
    got = abi.pull(h, A_CENTER, np.int32, 4)[:, :3]
    exp = want.export()
    assert np.array_equal(got, exp.vec("center"))
    assert abi.extent(h, A_DEADLIST) == 0
    abi.L.bt_world_destroy(h)


def test_grow_poke_spawn_emulated(oracle_lib):
    _grow_poke_spawn("emu", oracle_lib)


@pytest.mark.gpu
def test_grow_poke_spawn_cuda(oracle_lib):
    _grow_poke_spawn("cuda", oracle_lib)


def _full_hash_and_wires(backend, ref_lib):
    """bt_world_hash_full against the reference's restatement over its `ent`s on a level with sliders, wires, hold trees
    and buttons; then pending wires (three uWire + one uUnwire on one entity, in order) and a re-homed wire block"""
    R = ref_lib
    wid = R.new_world()
    R.load_sav("lights.sav")
    try:
        R.strip_handlers(drop_ticks=True, keep_push=0x3ff, keep_pushed=0x1ff, keep_pair=0x7f, keep_crush=0)
        R.kick(3, 120, 2)
        wa = R.export()
        abi = Abi(_so(backend))
        h = abi.world(wa, wa.n + 8, int(wa.desc.n_sliders) + 8, int(wa.desc.n_wires) + 8)
        assert abi.hash(h, full=True) == R.hash_full() and abi.hash(h) == R.hash()
        for t in range(12):
            R.step(1)
            abi.ck(abi.L.bt_world_step(h, 31))
            assert abi.hash(h) == R.hash(), t
            assert abi.hash(h, full=True) == R.hash_full(), t
        # e := the first entity that has wires; give it three more through the pending list, remove one of its own
        wo, wc = abi.pull(h, A_WIRE_OFF, np.int32), abi.pull(h, A_WIRE_CNT, np.int32)
        wt = abi.pull(h, A_WIRE_TO, np.int32)
        live = abi.pull(h, A_INLIST, np.uint8)
        e = next(i for i in range(len(wc)) if wc[i] > 0 and live[i])
        mine = [int(x) for x in wt[wo[e]: wo[e] + wc[e]]]
        others = [i for i in range(len(wc)) if live[i] and i not in mine and i != e][:3]
        targets = (C.c_int32 * len(mine))(*mine)
        abi.ck(abi.L.bt_world_set_wires(h, e, targets, len(mine), 4))      # too small a block: re-homed, same wires
        assert abi.pull(h, A_WIRE_CAP, np.int32)[e] >= len(mine) + 4
        abi.poke(h, [(OP_UWIRE, e, (others[0],)), (OP_UWIRE, e, (others[1],)), (OP_UUNWIRE, e, (mine[0],)),
                     (OP_UWIRE, e, (others[2],)), (OP_UWIRE, e, (others[0],)), (A_FULLFLUSH, e, (1,))])
        abi.ck(abi.L.bt_world_step(h, 2))                                  # BT_PHASE_PREP = flush1
        wo, wc = abi.pull(h, A_WIRE_OFF, np.int32), abi.pull(h, A_WIRE_CNT, np.int32)
        wt = abi.pull(h, A_WIRE_TO, np.int32)
        after = [int(x) for x in wt[wo[e]: wo[e] + wc[e]]]
        # flushMisc ent.cpp:694-712: adds in request order (no duplicates), then the removal pulls the LAST wire into the gap
        exp = mine + others
        k = exp.index(mine[0])
        exp[k] = exp[-1]
        exp.pop()
        assert after == exp, (after, exp)
        abi.L.bt_world_destroy(h)
    finally:
        R.free_world(wid)


def test_full_hash_and_pending_wires_emulated(ref_lib):
    _full_hash_and_wires("emu", ref_lib)


@pytest.mark.gpu
def test_full_hash_and_pending_wires_cuda(ref_lib):
    _full_hash_and_wires("cuda", ref_lib)


def test_poke_rejects_bad_input():
    abi = Abi(_so("emu"))
    wa, _ = load_scenario("stack512")
    h = abi.world(wa, wa.n + 4)
    buf = (Poke * 1)()
    buf[0].array, buf[0].index = A_CENTER, wa.n + 3       # beyond the slots in use
    assert abi.L.bt_world_poke(h, buf, 1) != 0
    buf[0].array, buf[0].index = A_SLIDER_OFF, 0          # read-only
    assert abi.L.bt_world_poke(h, buf, 1) != 0
    buf[0].array, buf[0].index = 77, 0                    # no such array
    assert abi.L.bt_world_poke(h, buf, 1) != 0
    first = C.c_int32()
    ns, wc = (C.c_int32 * 8)(), (C.c_int32 * 8)()
    assert abi.L.bt_world_grow(h, 8, ns, wc, 0, C.byref(first)) != 0   # capacity
    abi.L.bt_world_destroy(h)
