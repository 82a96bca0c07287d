"""The one-block-per-world step (k_tick_small, bittoss_b200/csrc/bt_small.cu: a whole doWholeStep of every small world in
ONE launch, loop control in shared memory) against the generic schedule (~40 launches per tick, option fused_small = 0)
and against the reference fixtures: per-tick state hash and full-state hash, host events in order, statistics, and the
downloaded state at the end.  Every world of <= 128 entities takes the fused path by default, so the handler labs, the
kicked shipped levels and the teacher-forced arena games of the other test files run through it too; this file pins
that both schedules are the same function."""
import numpy as np
import pytest

from util import SCENARIOS, golden_path, load_scenario, load_world

SMALL = [n for n in SCENARIOS if n in ("arena_kick", "diner_kick", "tower_kick", "forest_kick", "leggs_kick", "spider_kick",
                                       "spider_static")]
LABS = ["lab_teleport", "lab_bottles", "lab_seats", "lab_crush", "lab_logic", "lab_requests", "lab_rootorder", "lab_pilots", "lab_ticks2"]
SAME_STATS = ("ticks", "positives", "pushes", "serial_pushes", "crushes", "fumbles", "live_ents", "entity_ticks",
              "unsupported_events")


def _is_small(wa):
    return wa.n <= 128


def _pair(bt, wa):
    a, b = bt.World(wa), bt.World(wa)
    b.set_option("fused_small", 0)
    return a, b


def _same_arrays(x, y):
    assert x.n == y.n
    for k in x.a:
        assert np.array_equal(x.a[k], y.a[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", SMALL)
def test_cuda_fused_small_equals_generic_and_reference(bt, name):
    wa, ref = load_scenario(name)
    if not _is_small(wa):
        pytest.skip("%s holds %d entities" % (name, wa.n))
    f, g = _pair(bt, wa)
    for t in range(len(ref)):
        f.step(31)
        g.step(31)
        assert f.hash()[0] == g.hash()[0] == int(ref[t]), (name, t)
        assert f.hash_full()[0] == g.hash_full()[0], (name, t)
    sf, sg = f.stats(), g.stats()
    assert sf["kernel_launches"] < sg["kernel_launches"] // 3   # one launch per step (+ the hash kernels) against ~30
    for k in SAME_STATS + ("iterations",):
        assert sf[k] == sg[k], (name, k, sf[k], sg[k])
    # a full collisionBetter tie is counted whenever a proposal MEETS its equal, which depends on arrival order
    assert (sf["order_ties"] > 0) == (sg["order_ties"] > 0)
    assert sf["pushes"] > 0 or name == "spider_static"
    _same_arrays(f.download(), g.download())
    f.close()
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", LABS)
def test_cuda_fused_small_labs_events_and_requests(bt, name):
    """host requests between the steps, crush / push events out of them, the tick-handler phase: both schedules side by
    side (the labs' reference streams themselves are checked, through the fused path, by test_cuda_handler_labs)"""
    wa, extra = load_world(golden_path(name))
    assert _is_small(wa)
    script, ref = extra["script"], extra["ref_hashes"]
    ticks_phase = int(extra["mode"][0]) == 0
    f, g = _pair(bt, wa)
    for t in range(len(ref)):
        for (_, what, uid, other, a0, a1, a2) in script[script[:, 0] == t]:
            for w in (f, g):
                w.apply_requests([(int(what), 0, int(uid), int(other), (int(a0), int(a1), int(a2)))])
        for w in (f, g):
            if ticks_phase:
                w.tick(1)
            else:
                w.step(31)
        assert f.hash()[0] == g.hash()[0] == int(ref[t]), (name, t)
        assert f.hash_full()[0] == g.hash_full()[0], (name, t)
        assert f.drain_events() == g.drain_events(), (name, t)
    sf, sg = f.stats(), g.stats()
    for k in SAME_STATS:
        assert sf[k] == sg[k], (name, k, sf[k], sg[k])
    _same_arrays(f.download(), g.download())
    f.close()
    g.close()


@pytest.mark.gpu
def test_cuda_fused_small_batched_mixed_worlds(bt):
    """different levels, sizes (both block sizes: <= 64 and <= 128 slots) and labs in ONE device world, 300 copies:
    per-world hashes of both schedules, every tick; the generic schedule then takes over a world the fused one ran, and
    the other way round, mid-game"""
    names = ["arena_kick", "lab_crush", "diner_kick", "lab_seats", "tower_kick", "lab_teleport", "spider_kick"]
    ws = [load_world(golden_path(n))[0] for n in names]
    ws = [w for w in ws if _is_small(w)]
    assert len(ws) >= 5
    for batch in (ws, [w for w in ws if w.n <= 64] * 3):
        if not batch:
            continue
        wa = bt.concat_worlds(batch * (300 // len(batch)))
        f, g = _pair(bt, wa)
        for t in range(40):
            if t == 20:   # swap the schedules: everything a schedule keeps to itself (cached pairs, counters) must settle
                f.set_option("fused_small", 0)
                g.set_option("fused_small", 1)
            f.step(31)
            g.step(31)
            assert np.array_equal(f.hash(), g.hash()), t
        assert np.array_equal(f.hash_full(), g.hash_full())
        sf, sg = f.stats(), g.stats()
        for k in SAME_STATS:
            assert sf[k] == sg[k], (k, sf[k], sg[k])
        assert sorted(map(lambda e: tuple(sorted(e.items())), f.drain_events())) == \
            sorted(map(lambda e: tuple(sorted(e.items())), g.drain_events()))
        _same_arrays(f.download(), g.download())
        f.close()
        g.close()


@pytest.mark.gpu
def test_cuda_fused_small_partial_phases_and_host_deltas(bt):
    """steps that are not a whole tick (flushCtrls alone, prepPhysics alone, gravity without doPhysics: the pending d_vel is
    handed to the next launch), and host deltas / requests between fused ticks"""
    wa, _ = load_scenario("arena_kick")
    f, g = _pair(bt, wa)
    n = wa.n
    rng = np.random.RandomState(5)
    for t in range(30):
        dc = rng.randint(-3, 4, size=(n, 3)).astype(np.int32) if t % 7 == 3 else None
        for w in (f, g):
            if dc is not None and w.count()[0] == n:
                w.apply_host_deltas(dc, dc * 2)
            w.step(1)        # flushCtrls
            w.step(2)        # prepPhysics
            w.step(4)        # doGravity: leaves d_vel pending
            w.step(8 | 16)   # doPhysics + finishStep
        assert f.hash()[0] == g.hash()[0], t
        assert f.hash_full()[0] == g.hash_full()[0], t
    assert f.stats()["kernel_launches"] < g.stats()["kernel_launches"] // 2
    f.close()
    g.close()


@pytest.mark.gpu
def test_cuda_fused_small_pair_list_overflow_walks_the_boxes(bt):
    """a pile: every entity of the world within touching distance of most others, far more touching pairs than the
    shared-memory pair list holds (12 per entity) -- the narrow phase then walks the box triangle itself; the generic
    schedule (global list, no overflow) must see the same collisions, crushes included"""
    wa, _ = load_scenario("arena_kick")
    rng = np.random.RandomState(17)
    n = wa.n
    wa.a["center"][:] = rng.randint(-900, 901, size=3 * n).astype(np.int32)
    wa.a["old"][:] = wa.a["center"]
    wa.a["vel"][:] = rng.randint(-40, 41, size=3 * n).astype(np.int32)
    batch = bt.concat_worlds([wa] * 5)
    f, g = _pair(bt, batch)
    for t in range(25):
        f.step(31)
        g.step(31)
        assert np.array_equal(f.hash(), g.hash()), t
    sf, sg = f.stats(), g.stats()
    assert sf["positives"] == sg["positives"] and sf["positives"] > 20 * n   # a real pile
    for k in ("pushes", "crushes", "fumbles", "live_ents"):
        assert sf[k] == sg[k], (k, sf[k], sg[k])
    _same_arrays(f.download(), g.download())
    f.close()
    g.close()
