"""Kernel logic on the host: the per-thread bodies of bt_phys.cuh / bt_raycast.cuh, single-
stepped by tests/emu (same source the CUDA kernels compile), against the oracle."""
import numpy as np
import pytest

from util import SCENARIOS, load_scenario


@pytest.mark.parametrize("name", [s for s in SCENARIOS if s != "stack20k"])
@pytest.mark.parametrize("shuffle", [0, 11])
def test_kernel_bodies_match_oracle_and_reference(oracle_lib, emu, name, shuffle):
    wa, ref = load_scenario(name)
    ow = oracle_lib.create(wa)
    ew = emu.EmuWorld(wa, rebuild_every=3, shuffle_seed=shuffle)
    assert ew.hash() == ow.hash()
    for t in range(len(ref)):
        ow.step(31)
        ew.step(31)
        assert ew.hash() == ow.hash(), "tick %d" % t
        assert ew.hash()[0] == int(ref[t]), "reference golden, tick %d" % t
    so, se = ow.stats(), ew.stats()
    for k in ("iterations", "positives", "pushes", "crushes", "fumbles"):
        assert so[k] == se[k], k


def test_rebuild_period_does_not_change_results(emu):
    wa, ref = load_scenario("stack4k_pushall")
    a = emu.EmuWorld(wa, rebuild_every=1)
    b = emu.EmuWorld(wa, rebuild_every=50)
    for t in range(30):
        a.step(31)
        b.step(31)
        assert a.hash() == b.hash() == [int(ref[t])]


def test_batched_worlds_are_independent(oracle_lib, emu):
    """several worlds in one SoA (world id in the Morton key / node ranges) == each alone"""
    import bittoss_b200
    names = ["arena_kick", "diner_kick", "stack512", "arena_kick"]
    ws = [load_scenario(n)[0] for n in names]
    refs = [load_scenario(n)[1] for n in names]
    batch = bittoss_b200.concat_worlds(ws)
    ew = emu.EmuWorld(batch, rebuild_every=2, shuffle_seed=3)
    ow = oracle_lib.create(batch)
    for t in range(40):
        ew.step(31)
        ow.step(31)
        h = ew.hash()
        assert h == ow.hash()
        for g in range(len(names)):
            assert h[g] == int(refs[g][t]), (names[g], t)


def test_raycast_bodies_match_oracle(oracle_lib, emu):
    for name in ("lobby_kick", "stack512", "spider_static"):
        wa, _ = load_scenario(name)
        rng = np.random.RandomState(3)
        origin = rng.randint(0, wa.n, size=300).astype(np.int32)
        look = rng.randint(-64, 65, size=(300, 3)).astype(np.int32)
        look[::5, 1] = 0
        look[::9, 0] = 0
        look[np.abs(look).sum(1) == 0] = (0, 0, 64)
        want = oracle_lib.create(wa).selectcast(origin, look)
        got = emu.EmuWorld(wa).raycast(origin, look)
        assert np.array_equal(want, got), name


@pytest.mark.parametrize("name", ["stack4k_pushall", "spider_kick", "diner_kick"])
def test_pair_cache_modes_agree(emu, name):
    """cached candidate pairs (any margin, incl. 0 = constant invalidation) == traversal every iteration"""
    wa, ref = load_scenario(name)
    worlds = []
    for cache, margin in ((0, 16), (1, 0), (1, 16), (1, 200)):
        w = emu.EmuWorld(wa, rebuild_every=3, shuffle_seed=5)
        w.set_option("pair_cache", cache)
        w.set_option("cache_margin", margin)
        worlds.append(w)
    for t in range(len(ref)):
        for w in worlds:
            w.step(31)
            assert w.hash()[0] == int(ref[t]), t
    # a zero margin must have forced more pair rebuilds than a generous one
    assert worlds[1].pair_rebuilds() + worlds[1].pair_patches() >= worlds[3].pair_rebuilds() + worlds[3].pair_patches()


@pytest.mark.parametrize("name", ["stack4k_pushall", "stack2k_8high", "spider_kick", "diner_kick", "lobby_kick", "arena_kick"])
@pytest.mark.parametrize("rebuild_every", [2, 1000])
def test_incremental_pair_patches_agree(emu, name, rebuild_every):
    """patching the cached pair list for the entities that moved == rebuilding it from scratch"""
    if name not in SCENARIOS:
        pytest.skip("no such fixture")
    wa, ref = load_scenario(name)
    worlds = []
    for patch, margin, horizon in ((0, 8, 0), (1, 8, 0), (1, 0, 0), (1, 64, 0), (1, 8, 5), (0, 0, 3)):
        w = emu.EmuWorld(wa, rebuild_every=rebuild_every, shuffle_seed=11)
        w.set_option("pair_patch", patch)
        w.set_option("cache_margin", margin)
        w.set_option("cache_horizon", horizon)
        worlds.append(w)
    for t in range(len(ref)):
        for w in worlds:
            w.step(31)
            assert w.hash()[0] == int(ref[t]), t
    assert worlds[0].pair_patches() == 0
    if wa.n >= 1000:
        assert worlds[1].pair_patches() > 0


@pytest.mark.parametrize("name", ["arena_kick", "stack512"])
def test_small_world_pair_search_agrees_with_tree(emu, name):
    """worlds of <= 128 entities find their candidate pairs by brute force (no tree); same results"""
    if name == "stack512":
        import bittoss_b200
        from util import load_world, golden_path
        wa = bittoss_b200.concat_worlds([load_scenario("arena_kick")[0]] * 3)   # a batch of small worlds
        ref = load_scenario("arena_kick")[1]
    else:
        wa, ref = load_scenario(name)
    a = emu.EmuWorld(wa, rebuild_every=3, shuffle_seed=2)
    b = emu.EmuWorld(wa, rebuild_every=3, shuffle_seed=2)
    b.set_option("small_worlds", 0)
    for t in range(len(ref)):
        a.step(31)
        b.step(31)
        assert a.hash() == b.hash(), t
        assert a.hash()[0] == int(ref[t]), t
    assert a.pair_rebuilds() > 0
