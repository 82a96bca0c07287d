"""The CPU checkers: the plain-C restatement is pinned to outputs of the reference itself."""
import numpy as np
import pytest

from util import KEEP, SCENARIOS, load_scenario
from bittoss_b200.desc import state_hash

FAST = [s for s in SCENARIOS if s != "stack20k"]


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_oracle_matches_reference_golden(oracle_lib, name):
    """bit-exact per-tick hash stream vs the committed run of the unmodified reference"""
    wa, ref = load_scenario(name)
    ow = oracle_lib.create(wa)
    assert ow.hash()[0] == state_hash(wa)
    for t in range(len(ref)):
        ow.step(31)
        assert ow.hash()[0] == int(ref[t]), "%s diverges from the reference at tick %d" % (name, t)


@pytest.mark.parametrize("name", ["stack512", "spider_kick", "diner_kick", "arena_kick"])
def test_oracle_matches_reference_live_fullstate(oracle_lib, ref_lib, name):
    """with the reference compiled here: full exported state equality, not only the hash"""
    kind, args, ticks = SCENARIOS[name]
    R = ref_lib
    w = R.new_world()
    if kind == "stack":
        R.make_boxstack(*args)
    else:
        R.load_sav(args[0] + ".sav")
        R.strip_handlers(True, **KEEP)
        if args[2]:
            R.kick(args[1], args[2], 1)
    ow = oracle_lib.create(R.export())
    for t in range(ticks):
        R.step(1)
        ow.step(31)
        assert ow.hash()[0] == R.hash(), "tick %d" % t
    a, b = R.export().state_dict(), ow.export().state_dict()
    R.free_world(w)
    for k in a:
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k


def test_hash_python_restatement(oracle_lib):
    wa, _ = load_scenario("lobby_kick")
    assert oracle_lib.create(wa).hash()[0] == state_hash(wa)


def test_synth_generator_matches_reference_builder(ref_lib):
    """bt_synth_boxstack_fill (product host code) == the same recipe through the reference's initEnt"""
    import bittoss_b200
    wa = bittoss_b200.synth_boxstack(3000, 4, 777, 64, 60)
    w = ref_lib.new_world()
    ref_lib.make_boxstack(3000, 4, 777, 64, 60)
    ra = ref_lib.export()
    ref_lib.free_world(w)
    a, b = wa.state_dict(), ra.state_dict()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(wa.a["root_order"][: wa.n], ra.a["root_order"][: ra.n])


def test_oracle_raycast_matches_reference(oracle_lib, ref_lib):
    """selectCast: restatement vs raycast.cpp on lobby (whales, held ents, zero look components)"""
    R = ref_lib
    w = R.new_world()
    R.load_sav("lobby.sav")
    wa = R.export()
    rng = np.random.RandomState(5)
    n = wa.n
    origin = rng.randint(0, n, size=400).astype(np.int32)
    look = rng.randint(-64, 65, size=(400, 3)).astype(np.int32)
    look[::7, 0] = 0
    look[::11, 2] = 0
    look[look.sum(1) == 0] = (0, 0, 64)
    want = R.selectcast(origin, look)
    R.free_world(w)
    got = oracle_lib.create(wa).selectcast(origin, look)
    assert np.array_equal(want, got)
    assert (want >= 0).sum() > 50
