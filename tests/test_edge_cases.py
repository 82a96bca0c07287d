"""Empty, tiny and degenerate worlds — on the host-emulated kernel bodies and on CUDA."""
import numpy as np
import pytest

from bittoss_b200.desc import WorldArrays, state_hash


def tiny_world(n, radius=(400, 400, 400), spacing=0, held=False, boxed=True):
    """n cubes at the same place (spacing=0: everything overlaps everything) or in a row."""
    wa = WorldArrays(n, 0, 0, 1)
    wa.a["w_gamerules"][0] = 8  # EFFECT_GRAV
    for i in range(n):
        wa.a["uid"][i] = i
        wa.a["center"][3 * i: 3 * i + 3] = (i * spacing, 0, -i * 7)
        wa.a["old"][3 * i: 3 * i + 3] = (i * spacing, 0, -i * 7)
        wa.a["radius"][3 * i: 3 * i + 3] = radius
        wa.a["type_mask"][i] = wa.a["new_type_mask"][i] = 5 if boxed else 64
        wa.a["collide_mask"][i] = wa.a["new_collide_mask"][i] = 6 if boxed else 0
        wa.a["friction"][i] = 16
        wa.a["holder"][i] = -1 if (not held or i == 0) else 0
        wa.a["holdee_head"][i] = -1
        wa.a["holdee_next"][i] = -1
        wa.a["root_order"][i] = n - 1 - i
        wa.a["new_holder"][i] = -1
        wa.a["handlers"][10 * i] = 6  # WHOMOVES_WOOD
    if held and n > 1:  # entity 0 holds all others, list order newest first
        wa.a["holdee_head"][0] = n - 1
        for i in range(n - 1, 1, -1):
            wa.a["holdee_next"][i] = i - 1
        for i in range(1, n):
            wa.a["root_order"][i] = -1
        wa.a["root_order"][0] = 0
    return wa


CASES = {
    "empty": lambda: tiny_world(0),
    "single": lambda: tiny_world(1),
    "two_overlapping": lambda: tiny_world(2),
    "pile_of_40_same_place": lambda: tiny_world(40),           # mass crush / MOVE_BOTH storm
    "row_touching": lambda: tiny_world(30, spacing=800),       # exactly touching: not a collision
    "zero_radius": lambda: tiny_world(6, radius=(0, 0, 0)),
    "no_boxes_at_all": lambda: tiny_world(12, boxed=False),
    "one_tree_of_25": lambda: tiny_world(25, held=True, spacing=100),
}


@pytest.mark.parametrize("name", list(CASES))
def test_edge_world_kernel_bodies_vs_oracle(oracle_lib, emu, name):
    wa = CASES[name]()
    ow = oracle_lib.create(wa)
    ew = emu.EmuWorld(wa, rebuild_every=2, shuffle_seed=3)
    assert ew.hash() == ow.hash() == [state_hash(wa)]
    for t in range(40):
        ow.step(31)
        ew.step(31)
        assert ew.hash() == ow.hash(), "%s tick %d" % (name, t)


@pytest.mark.parametrize("name", list(CASES))
def test_edge_world_matches_reference(ref_lib, oracle_lib, name):
    """the same worlds through the reference's own initEnt/addEnt: oracle == reference"""
    import ctypes as C
    wa = CASES[name]()
    if wa.n == 0 or name == "one_tree_of_25":
        pytest.skip("needs entities / pickups the harness does not build")
    R = ref_lib
    wid = R.new_world()
    L = R.L
    L.btref_add_ent.argtypes = [C.POINTER(C.c_int32)] * 3 + [C.c_uint32, C.c_uint32, C.c_int]
    for i in range(wa.n):
        c = (C.c_int32 * 3)(*[int(x) for x in wa.a["center"][3 * i: 3 * i + 3]])
        v = (C.c_int32 * 3)(0, 0, 0)
        r = (C.c_int32 * 3)(*[int(x) for x in wa.a["radius"][3 * i: 3 * i + 3]])
        L.btref_add_ent(c, v, r, int(wa.a["type_mask"][i]), int(wa.a["collide_mask"][i]), 6)
    R.set_gamerules(8)
    ow = oracle_lib.create(R.export())
    for t in range(40):
        R.step(1)
        ow.step(31)
        assert ow.hash()[0] == R.hash(), "%s tick %d" % (name, t)
    R.free_world(wid)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_edge_world_cuda_vs_oracle(bt, oracle_lib, name):
    wa = CASES[name]()
    ow = oracle_lib.create(wa)
    w = bt.World(wa, max_ents=max(wa.n, 1))
    for t in range(40):
        ow.step(31)
        w.step(31)
        assert w.hash() == ow.hash(), "%s tick %d" % (name, t)
    a, b = ow.export().state_dict(), w.download().state_dict()
    for k in a:
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k
    if wa.n:
        origin = np.zeros(4, dtype=np.int32)
        look = np.array([[1, 0, 0], [0, 0, -1], [-3, 2, 1], [0, 64, 0]], dtype=np.int32)
        n_live = w.count()[0]
        if n_live:
            assert np.array_equal(w.raycast(origin, look), ow.selectcast(origin, look))
    w.close()


def falling_world(rules):
    """30 cubes and 10 terrain slabs with no ground under them: everything drops out of the world"""
    import ctypes as C
    wa = tiny_world(40, spacing=3000)
    wa.a["w_gamerules"][0] = 8 | rules
    for i in range(40):
        wa.a["center"][3 * i + 2] = 60000 + 40 * i
        wa.a["old"][3 * i + 2] = 60000 + 40 * i
        wa.a["center"][3 * i] = -58000 + 3000 * i
        wa.a["old"][3 * i] = -58000 + 3000 * i
        wa.a["vel"][3 * i] = 900 * (1 if i % 2 else -1)
        wa.a["old"][3 * i] -= 900 * (1 if i % 2 else -1)
        if i % 4 == 0:
            wa.a["type_mask"][i] = wa.a["new_type_mask"][i] = 2 | 16   # TERRAIN | WEIGHTLESS: bounces, does not die
            wa.a["collide_mask"][i] = wa.a["new_collide_mask"][i] = 0
    return wa


@pytest.mark.parametrize("rules", [16, 1, 2, 32, 1 | 16 | 32])
def test_world_rules_kernel_bodies_oracle_and_reference(ref_lib, oracle_lib, emu, rules):
    """BT_PHASE_WORLD_RULES: lava / crushtainers / bouncetainer, against the reference's effects.cpp"""
    import ctypes as C
    wa = falling_world(rules)
    R = ref_lib
    wid = R.new_world()
    R.L.btref_add_ent.argtypes = [C.POINTER(C.c_int32)] * 3 + [C.c_uint32, C.c_uint32, C.c_int]
    for i in range(wa.n):
        c = (C.c_int32 * 3)(*[int(x) for x in wa.a["center"][3 * i: 3 * i + 3]])
        v = (C.c_int32 * 3)(*[int(x) for x in wa.a["vel"][3 * i: 3 * i + 3]])
        r = (C.c_int32 * 3)(*[int(x) for x in wa.a["radius"][3 * i: 3 * i + 3]])
        R.L.btref_add_ent(c, v, r, int(wa.a["type_mask"][i]), int(wa.a["collide_mask"][i]), 6)
    R.set_gamerules(8 | rules)
    start = R.export()
    ow = oracle_lib.create(start)
    ew = emu.EmuWorld(start, rebuild_every=2, shuffle_seed=1)
    for t in range(120):
        R.step(2)
        ow.step(31 | 32)
        ew.step(31 | 32)
        assert ow.hash()[0] == R.hash(), "oracle != reference, tick %d" % t
        assert ew.hash() == ow.hash(), "kernel bodies != oracle, tick %d" % t
    assert R.count()[0] < wa.n   # something did fall out and die
    R.free_world(wid)


@pytest.mark.gpu
@pytest.mark.parametrize("rules", [16, 1 | 16 | 32])
def test_world_rules_cuda(bt, oracle_lib, rules):
    wa = falling_world(rules)
    ow = oracle_lib.create(wa)
    w = bt.World(wa)
    hs = w.tick(120, hashes=True)   # bt_world_tick adds the rules phase by itself
    for t in range(120):
        ow.step(31 | 32)
        assert int(hs[t, 0]) == ow.hash()[0], t
    assert w.count()[0] == ow.count()[0] < wa.n
    w.close()


def _ragged_batch():
    """worlds of very different sizes in one device world: the LBVH gives each its own subtree
    (a single-leaf world is just that leaf, a world without boxes has none)"""
    import bittoss_b200
    order = ["single", "pile_of_40_same_place", "empty", "no_boxes_at_all", "two_overlapping", "single",
             "one_tree_of_25", "row_touching"]
    return order, bittoss_b200.concat_worlds([CASES[k]() for k in order])


def test_ragged_batch_kernel_bodies_vs_oracle(oracle_lib, emu):
    order, batch = _ragged_batch()
    ow = oracle_lib.create(batch)
    ew = emu.EmuWorld(batch, rebuild_every=3, shuffle_seed=5)
    alone = [oracle_lib.create(CASES[k]()) for k in order]
    for t in range(30):
        ow.step(31)
        ew.step(31)
        for a in alone:
            a.step(31)
        assert ew.hash() == ow.hash() == [a.hash()[0] for a in alone], t
    rng = np.random.RandomState(1)
    n = ow.count()[0]
    origin = rng.randint(0, n, size=400).astype(np.int32)
    look = rng.randint(-64, 65, size=(400, 3)).astype(np.int32)
    look[np.abs(look).sum(1) == 0] = (0, 0, 64)
    assert np.array_equal(ow.selectcast(origin, look), ew.raycast(origin, look))


@pytest.mark.gpu
def test_ragged_batch_cuda_vs_oracle(bt, oracle_lib):
    order, batch = _ragged_batch()
    ow = oracle_lib.create(batch)
    w = bt.World(batch)
    for t in range(30):
        ow.step(31)
        w.step(31)
        assert [int(x) for x in w.hash()] == ow.hash(), t
    rng = np.random.RandomState(1)
    n = ow.count()[0]
    origin = rng.randint(0, n, size=400).astype(np.int32)
    look = rng.randint(-64, 65, size=(400, 3)).astype(np.int32)
    look[np.abs(look).sum(1) == 0] = (0, 0, 64)
    assert np.array_equal(ow.selectcast(origin, look), w.raycast(origin, look))
    w.close()


def _two_supporters():
    """a cube resting symmetrically on two cubes: its two candidate collisions tie in every
    collisionBetter criterion, the case `order_ties` counts"""
    wa = tiny_world(4)
    pos = [(0, 0, 2000), (-450, 0, 0), (450, 0, 0), (0, 0, -900)]    # a floor slab, two blocks side by side, one on top
    rad = [(3000, 3000, 300), (450, 450, 450), (450, 450, 450), (450, 450, 450)]
    for i, (p, r) in enumerate(zip(pos, rad)):
        wa.a["center"][3 * i: 3 * i + 3] = p
        wa.a["old"][3 * i: 3 * i + 3] = p
        wa.a["radius"][3 * i: 3 * i + 3] = r
    wa.a["center"][2] = wa.a["old"][2] = 750             # floor top at z = 450 = bottom of the two blocks
    wa.a["type_mask"][0] = wa.a["new_type_mask"][0] = 19  # terrain | heavy | weightless
    wa.a["collide_mask"][0] = wa.a["new_collide_mask"][0] = 0
    wa.a["handlers"][0] = 1                               # WHOMOVES_ME
    return wa


def test_full_ties_are_counted(oracle_lib, emu):
    wa = _two_supporters()
    ow = oracle_lib.create(wa)
    ew = emu.EmuWorld(wa, rebuild_every=2, shuffle_seed=1)
    for t in range(12):
        ow.step(31)
        ew.step(31)
        assert ew.hash() == ow.hash(), t
    assert ow.order_ties() > 0 and ew.order_ties() > 0


@pytest.mark.gpu
def test_full_ties_are_counted_cuda(bt, oracle_lib):
    wa = _two_supporters()
    ow = oracle_lib.create(wa)
    w = bt.World(wa)
    for t in range(12):
        ow.step(31)
        w.step(31)
        assert [int(x) for x in w.hash()] == ow.hash(), t
    assert w.stats()["order_ties"] > 0 and int(w.order_ties_per_world()[0]) > 0
    w.close()
