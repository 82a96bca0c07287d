"""Collision-path handlers no kicked shipped level triggers without players: PUSH_SEAT / TELEPORT /
PUDDLEJUMP / BOTTLE, PICKUP_SEAT, FUMBLE_SEAT(_OLD), crush handlers as host events, and the sparse
host-request channel (bt_world_apply_requests).  Fixtures = small hand-built worlds stepped by the
unmodified reference (tests/golden/make_handler_labs.py): per-tick hashes, the order the crush
handlers ran in, and the final sliders / wires / hold tree."""
import numpy as np
import pytest

from util import golden_path, load_world

LABS = ["lab_teleport", "lab_bottles", "lab_seats", "lab_crush", "lab_logic", "lab_requests", "lab_rootorder", "lab_pilots", "lab_ticks2"]
TICKS = 64  # BT_PHASE_TICKS


def _final_misc(extra):
    n, ns, nw = [int(x) for x in extra["final_n"]]
    uid = extra["final_uid"][:n]
    so, wo = extra["final_slider_off"], extra["final_wire_off"]
    out = {}
    for i in range(n):
        h = int(extra["final_holder"][i])
        out[int(uid[i])] = (int(uid[h]) if h >= 0 else -1,
                            [int(x) for x in extra["final_slider_v"][so[i]:so[i + 1]]],
                            [int(uid[x]) for x in extra["final_wire_to"][wo[i]:wo[i + 1]]],
                            [int(x) for x in extra["final_btn"][8 * i:8 * i + 8]])
    return out


def _run(name, make, request, step_hash, drain, misc):
    wa, extra = load_world(golden_path(name))
    script, ref = extra["script"], extra["ref_hashes"]
    # fixtures stepped with btref mode 0 ran the tick handlers too (doUpdates, ent.cpp:826-832)
    phases = 31 | (TICKS if int(extra["mode"][0]) == 0 else 0)
    w = make(wa)
    crush = []
    for t in range(len(ref)):
        for (_, what, uid, other, a0, a1, a2) in script[script[:, 0] == t]:
            request(w, int(what), int(uid), int(other), (int(a0), int(a1), int(a2)))
        assert step_hash(w, phases) == int(ref[t]), (name, t)
        crush += [(t, ev["me_uid"]) for ev in drain(w) if ev["kind"] == 3]
    assert crush == [tuple(int(x) for x in r) for r in extra["crush_order"]], name
    assert misc(w) == _final_misc(extra), name
    return w


@pytest.mark.parametrize("name", LABS)
@pytest.mark.parametrize("shuffle", [0, 7])
def test_emu_handler_labs(emu, name, shuffle):
    def step_hash(w, phases):
        w.step(phases)
        return w.hash()[0]

    w = _run(name, lambda wa: emu.EmuWorld(wa, rebuild_every=3, shuffle_seed=shuffle),
             lambda w, what, uid, other, a: w.request(what, uid, other, a),
             step_hash, lambda w: w.drain_events(), lambda w: w.export_misc())
    assert w.unsupported() == 0


def _desc_misc(st):
    a, n = st.a, st.n
    out = {}
    for i in range(n):
        h = int(a["holder"][i])
        out[int(a["uid"][i])] = (int(a["uid"][h]) if h >= 0 else -1,
                                 [int(x) for x in a["slider_v"][a["slider_off"][i]:a["slider_off"][i + 1]]],
                                 [int(a["uid"][x]) for x in a["wire_to"][a["wire_off"][i]:a["wire_off"][i + 1]]],
                                 [int(x) for x in a["btn"][8 * i:8 * i + 8]])
    return out


def _cuda_misc(w):
    return _desc_misc(w.download())


@pytest.mark.parametrize("name", LABS)
def test_oracle_handler_labs(oracle_lib, name):
    """the plain-C restatement against the same reference fixtures"""
    def step_hash(w, phases):
        w.step(phases)
        return w.hash()[0]

    _run(name, lambda wa: oracle_lib.create(wa),
         lambda w, what, uid, other, a: w.request(what, uid, other, a),
         step_hash, lambda w: [dict(kind=3, me_uid=u) for u in w.crush_log()], lambda w: _desc_misc(w.export()))


@pytest.mark.gpu
@pytest.mark.parametrize("name", LABS)
def test_cuda_handler_labs(bt, name):
    """strict upload (every handler id of the catalog has a device implementation or is an event)"""
    def step_hash(w, phases):
        if phases & TICKS:   # bt_world_tick adds the tick phase by itself when every handler has a device body
            return int(w.tick(1, hashes=True)[0, 0])
        w.step(phases)
        return int(w.hash()[0])

    w = _run(name, lambda wa: bt.World(wa),
             lambda w, what, uid, other, a: w.apply_requests([(what, 0, uid, other, a)]),
             step_hash, lambda w: w.drain_events(), _cuda_misc)
    assert w.stats()["unsupported_events"] == 0
    w.close()


@pytest.mark.gpu
def test_cuda_requests_reject_unknown_entity(bt):
    wa, _ = load_world(golden_path("lab_seats"))
    w = bt.World(wa)
    with pytest.raises(bt.BtError):
        w.apply_requests([(0, 0, 12345, -1, (0, 0, 0))])
    with pytest.raises(bt.BtError):
        w.apply_requests([(99, 0, 0, -1, (0, 0, 0))])
    w.close()


# ---- every shipped level with its FULL set of collision-path handlers (only tick handlers dropped) ----
ALL_LEVELS = ["arena", "blank", "ctf", "diner", "forest", "forts", "leggs", "lobby", "run", "ship", "soccer",
              "spacetrash", "speed", "spider", "tower", "w1-1"]   # lights: the reference itself is order-sensitive
                                                                   # under a kick, see test_oracle


def _full_catalog_level(R, level, seed, maxv):
    wid = R.new_world()
    R.load_sav(level if level == "w1-1" else level + ".sav")
    R.strip_handlers(True, keep_push=0xffffffff, keep_pushed=0xffffffff, keep_pair=0xffffffff, keep_crush=0xffffffff)
    R.lab_record_all_crushes()   # explode() would create entities on the host: record the calls instead
    R.lab_crush_log()
    R.kick(seed, maxv, 1)
    return wid


def _run_level_live(R, level, make, step_hash, drain, ticks=40):
    import os
    if not os.path.exists(os.path.join(os.path.dirname(R.L._name), "data", level if level == "w1-1" else level + ".sav")):
        pytest.skip("level not shipped")
    wid = _full_catalog_level(R, level, 77, 250)
    w = make(R.export())
    for t in range(ticks):
        R.step(1)
        assert step_hash(w) == R.hash(), (level, t)
        assert [ev["me_uid"] for ev in drain(w) if ev["kind"] == 3] == R.lab_crush_log(), (level, t)
    R.free_world(wid)
    return w


@pytest.mark.parametrize("level", ALL_LEVELS)
def test_emu_every_shipped_level_full_catalog(ref_lib, emu, level):
    def step_hash(w):
        w.step(31)
        return w.hash()[0]
    w = _run_level_live(ref_lib, level, lambda wa: emu.EmuWorld(wa, rebuild_every=4, shuffle_seed=3), step_hash,
                        lambda w: w.drain_events(), ticks=25)
    assert w.unsupported() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("level", ALL_LEVELS)
def test_cuda_every_shipped_level_full_catalog(ref_lib, bt, level):
    """strict upload: no handler id of any shipped level is refused, and 40 kicked ticks match the reference"""
    def step_hash(w):
        w.step(31)   # not tick(): that adds the world rules (crushtainer ...), which btref step mode 1 leaves out
        return int(w.hash()[0])
    w = _run_level_live(ref_lib, level, lambda wa: bt.World(wa), step_hash, lambda w: w.drain_events())
    assert w.stats()["unsupported_events"] == 0
    w.close()


# ---- shipped levels whose tick handlers all have device bodies: the FULL game tick on the device ----
TICK_LEVELS = ["tower", "forest", "speed", "blank", "lights", "arena", "lobby", "spider", "run", "soccer", "forts", "ctf", "diner", "leggs"]   # arena: + HELD_GUN (never held here), SCOREBOARD; lobby: + blink, jumper, loadblock, puddlejumper; spider: + its legs
EFFECT_BLOCKS = 4   # createDebris (effects.cpp:99-150) spawns platforms from gs->rand: host side, switched off here


def _device_tick_level(R, level, seed, maxv):
    wid = R.new_world()
    R.load_sav(level + ".sav")                     # all handlers stay, tick handlers included; no players
    R.set_gamerules(R.get_gamerules() & ~EFFECT_BLOCKS)
    R.lab_record_all_crushes()
    R.lab_crush_log()
    R.kick(seed, maxv, 1)
    return wid


def _run_tick_level(R, level, make, step_hash, maxv, ticks, reload=None):
    wid = _device_tick_level(R, level, 5, maxv)
    w = make(R.export())
    n0 = R.count()[0]
    reloads = 0
    for t in range(ticks):
        R.step(0)                                  # the reference's whole doWholeStep (main.cpp:1019-1097)
        if reload is not None and reload(w):
            # the scoreboard ended the round in the tick before (arena.sav without players: a tie after 70 ticks, everything
            # erased, scoreboard.cpp:81-86); requestReload is the host's: the reference has just reloaded the level
            if hasattr(w, "close"):
                w.close()
            w = make(R.export())
            reloads += 1
            continue
        assert step_hash(w) == R.hash(), (level, t)
    if reload is not None and level == "arena":
        assert reloads >= 1
    n1 = R.count()[0]
    R.free_world(wid)
    return w, n0, n1


# lights + a kick: the reference's own result depends on its list order there, no canonical answer
TICK_CASES = [(lv, v) for lv in TICK_LEVELS for v in (0, 120) if (lv, v) != ("lights", 120)]


@pytest.mark.parametrize("level,maxv", TICK_CASES)
def test_emu_levels_with_device_tick_handlers(ref_lib, emu, level, maxv):
    def step_hash(w):
        w.step(31 | TICKS | 32)
        return w.hash()[0]
    def round_over(w):   # kind 6 = a tick handler whose effect is the host's: TICK_SCOREBOARD asks for the level to be reloaded
        return any(ev["kind"] == 6 and ev["handler"] == 27 for ev in w.drain_events())
    w, n0, n1 = _run_tick_level(ref_lib, level, lambda wa: emu.EmuWorld(wa, rebuild_every=4, shuffle_seed=9),
                                step_hash, maxv, 60 if level != "arena" else 80, reload=round_over)
    assert w.unsupported() == 0


@pytest.mark.parametrize("level,maxv", TICK_CASES)
def test_oracle_levels_with_device_tick_handlers(ref_lib, oracle_lib, level, maxv):
    """the plain-C restatement of the same tick handlers against the reference's whole doWholeStep"""
    def step_hash(w):
        w.step(31 | TICKS | 32)
        return w.hash()[0]
    _run_tick_level(ref_lib, level, lambda wa: oracle_lib.create(wa), step_hash, maxv, 60)


@pytest.mark.gpu
@pytest.mark.parametrize("level,maxv", TICK_CASES)
def test_cuda_levels_with_device_tick_handlers(ref_lib, bt, level, maxv):
    """bt_world_tick alone = doWholeStep of the unmodified reference, tick handlers and world rules included"""
    def round_over(w):   # the device reports the tick handler whose effect is the host's: kind 6, TICK_SCOREBOARD
        return any(ev["kind"] == 6 and ev["handler"] == 27 for ev in w.drain_events())
    w, n0, n1 = _run_tick_level(ref_lib, level, lambda wa: bt.World(wa), lambda w: int(w.tick(1, hashes=True)[0, 0]),
                                maxv, 150, reload=round_over)
    st = w.stats()
    assert st["unsupported_events"] == 0 and st["live_ents"] == n1
    w.close()


@pytest.mark.gpu
def test_cuda_requests_across_batched_worlds(bt):
    """three labs as three worlds of ONE device world, each driven by its own script through one
    bt_world_apply_requests call per tick (requests interleaved across worlds): every world must
    follow its own fixture"""
    names = ["lab_seats", "lab_logic", "lab_requests", "lab_seats"]
    loaded = [load_world(golden_path(n)) for n in names]
    w = bt.World(bt.concat_worlds([wa for wa, _ in loaded]))
    ticks = min(len(ex["ref_hashes"]) for _, ex in loaded)
    for t in range(ticks):
        per_world = []
        for g, (_, ex) in enumerate(loaded):
            sc = ex["script"]
            per_world.append([(int(what), g, int(uid), int(other), (int(a0), int(a1), int(a2)))
                              for (_, what, uid, other, a0, a1, a2) in sc[sc[:, 0] == t]])
        reqs = []                              # round-robin merge: worlds interleaved, per-world order kept
        for i in range(max(len(p) for p in per_world)):
            reqs += [p[i] for p in per_world if i < len(p)]
        if reqs:
            w.apply_requests(reqs)
        hs = w.tick(1, hashes=True)[0]         # lab_logic needs the tick phase: bt_world_tick adds it (all worlds)
        for g, (_, ex) in enumerate(loaded):
            assert int(hs[g]) == int(ex["ref_hashes"][t]), (names[g], t)
    w.close()
