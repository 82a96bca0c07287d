"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden
hash streams.  Run on the B200 box: python -m pytest tests -m gpu"""
import numpy as np
import pytest

from util import SCENARIOS, load_scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_cuda_matches_oracle_and_reference_golden(bt, oracle_lib, name):
    wa, ref = load_scenario(name)
    ow = oracle_lib.create(wa)
    w = bt.World(wa)
    w.set_option("rebuild_every", 3)
    assert w.hash() == ow.hash()
    for t in range(len(ref)):
        ow.step(31)
        w.step(31)
        h = w.hash()
        assert h == ow.hash(), "%s: CUDA != oracle at tick %d" % (name, t)
        assert h[0] == int(ref[t]), "%s: CUDA != reference golden at tick %d" % (name, t)
    so, sg = ow.stats(), w.stats()
    for k in ("iterations", "positives", "pushes", "crushes", "fumbles"):
        assert so[k] == sg[k], k
    a, b = ow.export().state_dict(), w.download().state_dict()
    for k in a:
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k
    w.close()


def test_cuda_tick_hash_stream_api(bt):
    wa, ref = load_scenario("stack4k_pushall")
    w = bt.World(wa)
    hs = w.tick(len(ref), hashes=True)
    assert [int(x) for x in hs[:, 0]] == [int(x) for x in ref]
    w.close()


def test_cuda_batched_worlds(bt, oracle_lib):
    names = ["arena_kick", "diner_kick", "stack512", "arena_kick", "tower_kick"]
    ws = [load_scenario(n)[0] for n in names]
    refs = [load_scenario(n)[1] for n in names]
    w = bt.World(bt.concat_worlds(ws))
    for t in range(40):
        w.step(31)
        h = w.hash()
        for g in range(len(names)):
            assert h[g] == int(refs[g][t]), (names[g], t)
    w.close()


def test_cuda_100k_stack_vs_oracle(bt, oracle_lib):
    wa = bt.synth_boxstack(100000, 4, 777, 64, 60)
    ow = oracle_lib.create(wa)
    w = bt.World(wa)
    for t in range(12):
        ow.step(31)
        w.step(31)
        assert w.hash() == ow.hash(), "tick %d" % t
    w.close()


def _golden_stream(name):
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
    return z, [int(x) for x in z["hashes"]]


def test_cuda_1m_stack_vs_reference(bt):
    """BASELINE.json configs[2] AT ITS NAMED SIZE against the unmodified reference: 1 000 000 cubes + 6 400 ground tiles,
    the exact world bench.py times on rank 0, 75 ticks (50 settle + 5 warm-up + 20 timed) of per-tick hashes from
    tests/golden/stack1m_ref_hashes.npz (made by tests/golden/make_stack1m_ref.py from oracle/_ref: 15 minutes of CPU)"""
    z, ref = _golden_stream("stack1m_ref_hashes.npz")
    wa = bt.synth_boxstack(int(z["cubes"]), int(z["height"]), int(z["seed"]), int(z["push_every"]), int(z["push_max"]))
    assert wa.n == int(z["entities"])
    w = bt.World(wa)
    got = [int(x) for x in w.tick(75, hashes=True)[:, 0]]
    assert got == ref[:75], "first differing tick: %d" % next(i for i in range(75) if got[i] != ref[i])
    assert w.stats()["live_ents"] == int(z["live"][74])
    w.close()


def test_cuda_1m_pushall_vs_reference(bt):
    """the same world with the top cube of EVERY column kicked (250 000 sliding cubes: pair lists are patched and
    rebuilt while they move), against the reference's stream for it"""
    import os
    if not os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stack1m_pushall_ref_hashes.npz")):
        pytest.skip("fixture not generated")
    z, ref = _golden_stream("stack1m_pushall_ref_hashes.npz")
    wa = bt.synth_boxstack(int(z["cubes"]), int(z["height"]), int(z["seed"]), int(z["push_every"]), int(z["push_max"]))
    w = bt.World(wa)
    n = len(ref)
    got = [int(x) for x in w.tick(n, hashes=True)[:, 0]]
    assert got == ref, "first differing tick: %d" % next(i for i in range(n) if got[i] != ref[i])
    st = w.stats()
    assert st["crushes"] == wa.n - int(z["live"][-1]) and st["crushes"] > 1000   # the kicked cubes really collide
    w.close()


def test_cuda_1m_stack_properties(bt):
    """full BASELINE size: no oracle run (minutes on a CPU); size-independent properties:
    determinism run-to-run, independence from the tree rebuild period, entity conservation"""
    wa = bt.synth_boxstack(1000000, 4, 777, 64, 60)
    hs = []
    for period in (1, 8, 8):
        w = bt.World(wa)
        w.set_option("rebuild_every", period)
        hs.append([int(x) for x in w.tick(12, hashes=True)[:, 0]])
        st = w.stats()
        assert st["live_ents"] + st["crushes"] == wa.n
        w.close()
    assert hs[0] == hs[1] == hs[2]
    assert len(set(hs[0])) > 4


def test_cuda_raycast_matches_oracle(bt, oracle_lib):
    for name in ("lobby_kick", "stack4k_pushall", "spider_static"):
        wa, _ = load_scenario(name)
        rng = np.random.RandomState(3)
        origin = rng.randint(0, wa.n, size=2000).astype(np.int32)
        look = rng.randint(-64, 65, size=(2000, 3)).astype(np.int32)
        look[::5, 1] = 0
        look[::9, 0] = 0
        look[np.abs(look).sum(1) == 0] = (0, 0, 64)
        want = oracle_lib.create(wa).selectcast(origin, look)
        w = bt.World(wa)
        got = w.raycast(origin, look)
        w.close()
        assert np.array_equal(want, got), name


def test_cuda_raycast_after_ticks_skips_dead(bt, oracle_lib):
    wa, _ = load_scenario("spider_static")
    ow = oracle_lib.create(wa)
    w = bt.World(wa)
    for _ in range(20):
        ow.step(31)
        w.step(31)
    n = w.count()[0]
    assert n == ow.count()[0] and n < wa.n
    rng = np.random.RandomState(9)
    origin = rng.randint(0, n, size=1000).astype(np.int32)
    look = rng.randint(-64, 65, size=(1000, 3)).astype(np.int32)
    look[np.abs(look).sum(1) == 0] = (0, 0, 64)
    assert np.array_equal(ow.selectcast(origin, look), w.raycast(origin, look))
    w.close()


def test_unsupported_handler_is_refused(bt):
    wa, _ = load_scenario("arena_kick")
    wa.a["handlers"][4] = 10  # push id PUSH_NUM on entity 0: not in the catalog (handlerRegistrar.h:163-175)
    with pytest.raises(bt.BtError):
        bt.World(wa)
    wa.a["handlers"][4] = 0
    wa.a["handlers"][6] = 6  # onFumble = ENTPAIR_reuse1: no registered handler
    with pytest.raises(bt.BtError):
        bt.World(wa)


@pytest.mark.parametrize("name", ["stack4k_pushall", "spider_kick", "stack20k"])
def test_cuda_pair_cache_modes_agree(bt, name):
    """cached candidate pairs (k_pairs + k_narrow, any margin) == traversal every iteration (k_collide)"""
    wa, ref = load_scenario(name)
    for cache, margin, patch in ((0, 16, 1), (1, 0, 1), (1, 32, 1), (1, 300, 1), (1, 0, 0), (1, 32, 0)):
        w = bt.World(wa)
        w.set_option("pair_cache", cache)
        w.set_option("cache_margin", margin)
        w.set_option("pair_patch", patch)
        hs = w.tick(len(ref), hashes=True)
        assert [int(x) for x in hs[:, 0]] == [int(x) for x in ref], (cache, margin, patch)
        if not patch:
            assert w.stats()["pair_patches"] == 0
        w.close()


def test_cuda_raycast_100k_world_vs_oracle(bt, oracle_lib):
    """config-5 style: rays from random entities of a settled 100k-cube world, subsample vs brute force"""
    wa = bt.synth_boxstack(100000, 4, 777, 64, 60)
    w = bt.World(wa)
    w.tick(30)
    st = w.download()
    rng = np.random.RandomState(11)
    origin = rng.randint(0, st.n, size=2048).astype(np.int32)
    look = rng.randint(-64, 65, size=(2048, 3)).astype(np.int32)
    look[np.abs(look).sum(1) == 0] = (0, 0, 64)
    got = w.raycast(origin, look)
    want = oracle_lib.create(st).selectcast(origin, look)
    assert np.array_equal(got, want)
    assert (got >= 0).mean() > 0.3
    w.close()


def test_cuda_host_exchange_roundtrip(bt, oracle_lib):
    """bt_world_apply_host_deltas == uVel/uCenter on the host side; bt_world_read_state == download"""
    wa, _ = load_scenario("tower_kick")
    rng = np.random.RandomState(2)
    dc = rng.randint(-20, 21, size=(wa.n, 3)).astype(np.int32)
    dv = rng.randint(-50, 51, size=(wa.n, 3)).astype(np.int32)
    w = bt.World(wa)
    w.apply_host_deltas(dc, dv)
    wa.a["d_center"][: 3 * wa.n] += dc.reshape(-1)
    wa.a["d_vel"][: 3 * wa.n] += dv.reshape(-1)
    ow = oracle_lib.create(wa)
    for t in range(20):
        w.step(31)
        ow.step(31)
        assert w.hash() == ow.hash(), t
    n = w.count()[0]
    c = np.zeros((n, 3), dtype=np.int32)
    v = np.zeros((n, 3), dtype=np.int32)
    w.read_state(c, v)
    st = w.download()
    assert np.array_equal(c, st.vec("center")) and np.array_equal(v, st.vec("vel"))
    w.close()


def test_cuda_schedule_variants_agree_at_scale(bt, oracle_lib):
    """200k cubes, 90 ticks: run-ahead on/off x pair cache on/off x rebuild period must give one
    hash stream (grids larger than one wave, many blocks, stand-downs), and it must be the oracle's"""
    wa = bt.synth_boxstack(200000, 4, 777, 64, 60)
    streams = []
    patches = []
    for run_ahead, cache, period, patch, pdl in ((1, 1, 16, 1, -1), (0, 1, 16, 1, 0), (1, 1, 3, 1, 1), (0, 0, 5, 1, 3),
                                                 (1, 1, 16, 0, 1), (1, 1, 1000, 1, 0), (0, 1, 4, 1, 2)):
        w = bt.World(wa)
        w.set_option("pdl", pdl)   # dependent launches: auto / off / short kernels only / all kernels
        w.set_option("run_ahead", run_ahead)
        w.set_option("pair_cache", cache)
        w.set_option("rebuild_every", period)
        w.set_option("pair_patch", patch)
        streams.append([int(x) for x in w.tick(90, hashes=True)[:, 0]])
        patches.append(w.stats()["pair_patches"])
        w.close()
    assert all(s == streams[0] for s in streams[1:])
    assert patches[0] > 0 and patches[4] == 0  # the incremental path was really exercised
    ow = oracle_lib.create(wa)
    for t in range(10):
        ow.step(31)
        assert ow.hash()[0] == streams[0][t], t


def test_cuda_dup_and_rollback(bt):
    """bt_world_dup / bt_world_assign = dup() (ent.cpp:1135-1200): the copy evolves identically,
    and restoring a snapshot replays the same hashes (the client's rollback loop, main.cpp:1265-1280)"""
    wa, ref = load_scenario("spider_kick")
    w = bt.World(wa)
    w.tick(7)
    snap = w.dup()
    a = [int(x) for x in w.tick(15, hashes=True)[:, 0]]
    b = [int(x) for x in snap.tick(15, hashes=True)[:, 0]]
    assert a == b == [int(x) for x in ref[7:22]]
    snap2 = w.dup()          # state at tick 22
    w.tick(5)
    w.assign(snap2)          # roll back to tick 22
    c = [int(x) for x in w.tick(10, hashes=True)[:, 0]]
    assert c == [int(x) for x in ref[22:32]]
    big = bt.World(bt.synth_boxstack(50000))
    big.tick(20)
    twin = big.dup()
    assert [int(x) for x in big.tick(10, hashes=True)[:, 0]] == [int(x) for x in twin.tick(10, hashes=True)[:, 0]]
    for x in (w, snap, snap2, big, twin):
        x.close()


def test_cuda_small_world_pair_search_agrees_with_tree(bt):
    """small-world mode (brute-force pairs per world, one block each) vs the LBVH path, single and batched"""
    wa1, ref = load_scenario("arena_kick")
    for wa in (wa1, bt.concat_worlds([wa1] * 40)):
        streams = []
        for small in (1, 0):
            w = bt.World(wa)
            w.set_option("small_worlds", small)
            hs = []
            for _ in range(len(ref)):
                w.step(31)   # the fixture was stepped without the world rules bt_world_tick would add
                hs.append(w.hash())
            streams.append(np.array([[int(x) for x in h] for h in hs], dtype=np.uint64))
            st = w.stats()
            assert (st["tree_builds"] == 0) == bool(small)
            w.close()
        assert np.array_equal(streams[0], streams[1])
        assert [int(x) for x in streams[0][:, 0]] == [int(x) for x in ref]


def test_cuda_raycast_large_batch_sorted_equals_unsorted_and_oracle(bt, oracle_lib):
    """batches of >= 4096 rays: handed out dynamically to the warps or one per thread, in the caller's order or in the order of
    their origin's leaf position (radix-sorted on the device); the winners come back by ray index and equal the oracle's
    brute force in every mode"""
    wa, _ = load_scenario("stack4k_pushall")
    rng = np.random.RandomState(11)
    n = 20000
    look = rng.randint(-64, 65, size=(n, 3)).astype(np.int32)
    look[np.abs(look).sum(1) == 0] = (0, 0, 64)
    ow = oracle_lib.create(wa)
    w = bt.World(wa)
    for _ in range(6):
        ow.step(31)
        w.step(31)
    live = w.count()[0]
    assert live == ow.count()[0] and live < wa.n   # some cubes were crushed: ranks are live ranks
    origin = rng.randint(0, live, size=n).astype(np.int32)
    origin[::97] = -1                      # not a live entity: no caster, no hit
    want = ow.selectcast(np.where(origin < 0, 0, origin), look)
    want[origin < 0] = -1
    got_plain = w.raycast(origin, look)        # rays handed out dynamically to the warps (k_raycast_dyn, the default)
    w.set_option("sort_rays", 1)
    got_sorted = w.raycast(origin, look)
    w.set_option("ray_dynamic", 0)             # one ray per thread, fixed (k_raycast)
    got_static_sorted = w.raycast(origin, look)
    w.set_option("sort_rays", 0)
    got_static = w.raycast(origin, look)
    w.set_option("ray_dynamic", 1)
    w.set_option("ray_blocks", 3)              # far fewer warps than chunks of rays: every warp refills many times
    got_few = w.raycast(origin, look)
    w.close()
    for got in (got_plain, got_sorted, got_static_sorted, got_static, got_few):
        assert np.array_equal(got, want)


def test_cuda_raycast_matches_reference_selectcast(bt, ref_lib):
    """against the REFERENCE's own selectCast (raycast.cpp:16-57), not the port: the state the device has reached after
    kicked ticks is downloaded and rebuilt as `ent`s through the reference's initEnt / pickupNoHandlers
    (btref_build_from_desc), hold trees included (a caster's own tree is skipped, raycast.cpp:34)"""
    for name, ticks in (("lobby_kick", 15), ("spider_kick", 10), ("stack4k_pushall", 8)):
        wa, _ = load_scenario(name)
        w = bt.World(wa)
        for _ in range(ticks):
            w.step(31)
        state = w.download()
        rng = np.random.RandomState(5)
        origin = rng.randint(0, state.n, size=1500).astype(np.int32)
        look = rng.randint(-64, 65, size=(1500, 3)).astype(np.int32)
        look[::7, 2] = 0
        look[np.abs(look).sum(1) == 0] = (0, 0, 64)
        got = w.raycast(origin, look)
        w.close()
        wid = ref_lib.new_world()
        try:
            assert ref_lib.build_from(state) == state.n
            want = ref_lib.selectcast(origin, look)
        finally:
            ref_lib.free_world(wid)
        assert np.array_equal(got, want), name@pytest.mark.gpu
def test_cuda_pipelined_host_exchange_equals_serial(bt):
    """bt_world_apply_host_deltas_async / bt_world_read_state_async (copies on their own streams, double-buffered, overlapping
    the ticks) against the serial calls: same states read back every tick, same hashes, also while entities die"""
    import torch
    wa, _ = load_scenario("stack4k_pushall")
    a, b = bt.World(wa), bt.World(wa)
    rng = np.random.RandomState(3)
    n = wa.n
    pin = lambda: torch.zeros((n, 3), dtype=torch.int32).pin_memory().numpy()
    bufs = [[pin() for _ in range(4)] for _ in range(2)]
    sc, sv = np.zeros((n, 3), np.int32), np.zeros((n, 3), np.int32)
    deltas = [(rng.randint(-2, 3, size=(n, 3)).astype(np.int32), rng.randint(-1, 2, size=(n, 3)).astype(np.int32)) for _ in range(14)]
    live = n
    got = []
    tickets = [None, None]
    bufs[0][0][:] = deltas[0][0]; bufs[0][1][:] = deltas[0][1]
    b.apply_host_deltas_async(bufs[0][0][:live], bufs[0][1][:live])
    for t in range(14):
        k = t & 1
        # serial world
        la = a.live_ents()
        a.apply_host_deltas(np.ascontiguousarray(deltas[t][0][:la]), np.ascontiguousarray(deltas[t][1][:la]))
        a.step(31)
        la = a.live_ents()
        a.read_state(sc[:la], sv[:la])
        want = (sc[:la].copy(), sv[:la].copy(), a.hash()[0])
        # pipelined world
        b.step(31)
        lb = b.live_ents()
        assert lb == la
        tickets[k] = b.read_state_async(bufs[k][2][:lb], bufs[k][3][:lb])
        if t + 1 < 14:
            bufs[k ^ 1][0][:] = deltas[t + 1][0]; bufs[k ^ 1][1][:] = deltas[t + 1][1]
            b.apply_host_deltas_async(bufs[k ^ 1][0][:lb], bufs[k ^ 1][1][:lb])
        hb = b.hash()[0]
        b.read_state_wait(tickets[k])
        assert hb == want[2], t
        assert np.array_equal(bufs[k][2][:lb], want[0]) and np.array_equal(bufs[k][3][:lb], want[1]), t
    assert a.live_ents() < n   # cubes were crushed on the way: the arrays shrank with the world
    a.close()
    b.close()


@pytest.mark.gpu
def test_cuda_schedule_variants_agree_at_scale(bt, oracle_lib):
    """200k cubes, 90 ticks: run-ahead on/off x pair cache on/off x rebuild period must give one
    hash stream (grids larger than one wave, many blocks, stand-downs), and it must be the oracle's"""
    wa = bt.synth_boxstack(200000, 4, 777, 64, 60)
    streams = []
    patches = []
    for run_ahead, cache, period, patch, pdl in ((1, 1, 16, 1, -1), (0, 1, 16, 1, 0), (1, 1, 3, 1, 1), (0, 0, 5, 1, 3),
                                                 (1, 1, 16, 0, 1), (1, 1, 1000, 1, 0), (0, 1, 4, 1, 2)):
        w = bt.World(wa)
        w.set_option("pdl", pdl)   # dependent launches: auto / off / short kernels only / all kernels
        w.set_option("run_ahead", run_ahead)
        w.set_option("pair_cache", cache)
        w.set_option("rebuild_every", period)
        w.set_option("pair_patch", patch)
        streams.append([int(x) for x in w.tick(90, hashes=True)[:, 0]])
        patches.append(w.stats()["pair_patches"])
        w.close()
    assert all(s == streams[0] for s in streams[1:])
    assert patches[0] > 0 and patches[4] == 0  # the incremental path was really exercised
    ow = oracle_lib.create(wa)
    for t in range(10):
        ow.step(31)
        assert ow.hash()[0] == streams[0][t], t


def test_cuda_dup_and_rollback(bt):
    """bt_world_dup / bt_world_assign = dup() (ent.cpp:1135-1200): the copy evolves identically,
    and restoring a snapshot replays the same hashes (the client's rollback loop, main.cpp:1265-1280)"""
    wa, ref = load_scenario("spider_kick")
    w = bt.World(wa)
    w.tick(7)
    snap = w.dup()
    a = [int(x) for x in w.tick(15, hashes=True)[:, 0]]
    b = [int(x) for x in snap.tick(15, hashes=True)[:, 0]]
    assert a == b == [int(x) for x in ref[7:22]]
    snap2 = w.dup()          # state at tick 22
    w.tick(5)
    w.assign(snap2)          # roll back to tick 22
    c = [int(x) for x in w.tick(10, hashes=True)[:, 0]]
    assert c == [int(x) for x in ref[22:32]]
    big = bt.World(bt.synth_boxstack(50000))
    big.tick(20)
    twin = big.dup()
    assert [int(x) for x in big.tick(10, hashes=True)[:, 0]] == [int(x) for x in twin.tick(10, hashes=True)[:, 0]]
    for x in (w, snap, snap2, big, twin):
        x.close()


def test_cuda_small_world_pair_search_agrees_with_tree(bt):
    """small-world mode (brute-force pairs per world, one block each) vs the LBVH path, single and batched"""
    wa1, ref = load_scenario("arena_kick")
    for wa in (wa1, bt.concat_worlds([wa1] * 40)):
        streams = []
        for small in (1, 0):
            w = bt.World(wa)
            w.set_option("small_worlds", small)
            hs = []
            for _ in range(len(ref)):
                w.step(31)   # the fixture was stepped without the world rules bt_world_tick would add
                hs.append(w.hash())
            streams.append(np.array([[int(x) for x in h] for h in hs], dtype=np.uint64))
            st = w.stats()
            assert (st["tree_builds"] == 0) == bool(small)
            w.close()
        assert np.array_equal(streams[0], streams[1])
        assert [int(x) for x in streams[0][:, 0]] == [int(x) for x in ref]


def test_cuda_raycast_large_batch_sorted_equals_unsorted_and_oracle(bt, oracle_lib):
    """batches of >= 4096 rays: handed out dynamically to the warps or one per thread, in the caller's order or in the order of
    their origin's leaf position (radix-sorted on the device); the winners come back by ray index and equal the oracle's
    brute force in every mode"""
    wa, _ = load_scenario("stack4k_pushall")
    rng = np.random.RandomState(11)
    n = 20000
    look = rng.randint(-64, 65, size=(n, 3)).astype(np.int32)
    look[np.abs(look).sum(1) == 0] = (0, 0, 64)
    ow = oracle_lib.create(wa)
    w = bt.World(wa)
    for _ in range(6):
        ow.step(31)
        w.step(31)
    live = w.count()[0]
    assert live == ow.count()[0] and live < wa.n   # some cubes were crushed: ranks are live ranks
    origin = rng.randint(0, live, size=n).astype(np.int32)
    origin[::97] = -1                      # not a live entity: no caster, no hit
    want = ow.selectcast(np.where(origin < 0, 0, origin), look)
    want[origin < 0] = -1
    got_plain = w.raycast(origin, look)        # rays handed out dynamically to the warps (k_raycast_dyn, the default)
    w.set_option("sort_rays", 1)
    got_sorted = w.raycast(origin, look)
    w.set_option("ray_dynamic", 0)             # one ray per thread, fixed (k_raycast)
    got_static_sorted = w.raycast(origin, look)
    w.set_option("sort_rays", 0)
    got_static = w.raycast(origin, look)
    w.set_option("ray_dynamic", 1)
    w.set_option("ray_blocks", 3)              # far fewer warps than chunks of rays: every warp refills many times
    got_few = w.raycast(origin, look)
    w.close()
    for got in (got_plain, got_sorted, got_static_sorted, got_static, got_few):
        assert np.array_equal(got, want)


def test_cuda_raycast_matches_reference_selectcast(bt, ref_lib):
    """against the REFERENCE's own selectCast (raycast.cpp:16-57), not the port: the state the device has reached after
    kicked ticks is downloaded and rebuilt as `ent`s through the reference's initEnt / pickupNoHandlers
    (btref_build_from_desc), hold trees included (a caster's own tree is skipped, raycast.cpp:34)"""
    for name, ticks in (("lobby_kick", 15), ("spider_kick", 10), ("stack4k_pushall", 8)):
        wa, _ = load_scenario(name)
        w = bt.World(wa)
        for _ in range(ticks):
            w.step(31)
        state = w.download()
        rng = np.random.RandomState(5)
        origin = rng.randint(0, state.n, size=1500).astype(np.int32)
        look = rng.randint(-64, 65, size=(1500, 3)).astype(np.int32)
        look[::7, 2] = 0
        look[np.abs(look).sum(1) == 0] = (0, 0, 64)
        got = w.raycast(origin, look)
        w.close()
        wid = ref_lib.new_world()
        try:
            assert ref_lib.build_from(state) == state.n
            want = ref_lib.selectcast(origin, look)
        finally:
            ref_lib.free_world(wid)
        assert np.array_equal(got, want), name
