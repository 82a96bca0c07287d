"""Free-running hybrid: the reference's own host code (tick handlers, players, spawners, crush handlers,
serialize: compiled unmodified into host/_build/libbthost.so) with its physics on the device through
host/bt_glue.cpp, against the pure reference (oracle/_ref) on the same seeded inputs.

Nothing is re-uploaded: the world is uploaded once (bt_glue_attach); afterwards only what host code changed
goes to the device (sparse writes, new entities) and what the device changed comes back into the `ent`s.  Every
tick three things must agree with the reference: the state hash and the full-state hash computed (a) from the
host's `ent`s after the write-back and (b) on the device from its arrays; every 100 ticks the bytes the
reference's own serialize() (serialize.cpp:189-212) writes from the hybrid's gamestate must equal the pure run's.

`-m "not gpu"`: the device side is the kernel-body emulator (tests/emu, same C ABI subset), short runs.
`-m gpu`: libbittoss_b200.so on cuda:0, the tick counts BASELINE.json names (configs[0], [1], [3]).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pybind  # noqa: E402
from bittoss_b200 import hostlib  # noqa: E402

needs_ref = pytest.mark.skipif(not (pybind.have_ref() and hostlib.available()),
                               reason="oracle/_ref or host/_build not built (reference tree absent at build time)")

# (level, player data, spacetrash spawner on)
LOBBY = ("lobby.sav", (0, 0, 0, 0), 0)
SPACETRASH = ("spacetrash.sav", (), 1)
ARENA = ("arena.sav", (1, 2), 0)
OTHER_LEVELS = ["ctf", "forts", "diner", "forest", "leggs", "lights", "run", "ship", "soccer", "speed", "spider", "tower",
                "blank"]


def _host(backend):
    if backend == "emu":
        import emubind
        emubind.build_emu()
        so = emubind.EMU_SO
    else:
        so = None  # the CUDA library of this repository
    try:
        return hostlib.HeadlessHost(so)
    except RuntimeError as e:
        if "already bound" in str(e):
            pytest.skip("the headless host of this process is bound to the other backend (run -m gpu and -m 'not gpu' separately)")
        raise


def free_run(backend, level, players, spacetrash, ticks, serialize_every=100, input_key=0):
    R = pybind.RefLib()
    H = _host(backend)
    rw, hw = R.new_world(), H.new_world()
    try:
        R.load_sav(level, players)
        R.set_options(spacetrash, input_key)
        H.load_sav(level, players)
        H.set_options(spacetrash, input_key)
        H.attach(0, 65536)
        assert R.serialize() == H.serialize()
        for t in range(ticks):
            with pybind.stdout_to_stderr():
                R.select(rw)
                R.step(0)       # the reference's whole doWholeStep, CPU
                H.select(hw)
                H.step()        # the same host code, physics on the device
            ref = (R.hash(), R.hash_full())
            host, dev = H.hash(), H.device_hash()
            assert host == ref, "tick %d: host `ent`s after write-back differ from the reference (%d vs %d entities)" % (
                t, H.count(), R.count()[0])
            assert dev == ref, "tick %d: device arrays differ from the reference" % t
            if (t + 1) % serialize_every == 0:
                assert H.serialize() == R.serialize(), "tick %d: serialize() bytes differ" % t
        assert H.serialize() == R.serialize()
        c = H.counters()
        return c, H.count()
    finally:
        R.free_world(rw)
        H.free_world(hw)


# ---------------------------------------------------------------- CPU: emulator backend, short runs

@needs_ref
@pytest.mark.parametrize("case,ticks", [(LOBBY, 250), (ARENA, 300), (SPACETRASH, 1200)], ids=["lobby_4players", "arena_2players", "spacetrash_spawner"])
def test_free_running_emulated(case, ticks):
    c, n = free_run("emu", case[0], case[1], case[2], ticks)
    assert c["pulls"] >= ticks and c["spawned"] > 0


@needs_ref
@pytest.mark.parametrize("level", OTHER_LEVELS)
def test_free_running_every_level_emulated(level):
    free_run("emu", level + ".sav", (1, 2), 0, 120, serialize_every=40)


def _crush_lab_sav(name):
    """a world whose crush handlers really run (flag_crush -> explode(): 64 new entities; platform_crush ->
    uSlider on the wired entities), built through the reference and written as a .sav both sides load"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_handler_labs as labs
    R = pybind.RefLib()
    w = R.new_world()
    R.lab_real_crush(True)
    try:
        floor = labs.pad(R, 0, 0, r=(6000, 6000, 300), z=400)
        victims = []
        for i in range(4):
            e = labs.cube(R, -3000 + 2000 * i, 0, -300, r=400)
            R.lab_set_handlers(e, crush=1 + (i & 1))          # CRUSH_FLAG, CRUSH_PLATFORM alternating
            victims.append(e)
        watcher = labs.cube(R, 0, 4000, -300, r=300)           # what platform_crush writes to (player.h:1-2 sliders 8, 9)
        R.lab_set_sliders(watcher, [0] * 12)
        R.lab_add_wire(victims[1], watcher)
        R.lab_add_wire(victims[3], watcher)
        R.add_ent((0, 0, -1300), (0, 0, 200), (5000, 5000, 200), labs.GROUND, 0, labs.WM_ME)  # an unstoppable slab
        R.set_gamerules(8)  # EFFECT_GRAV
        data = R.serialize()
    finally:
        R.lab_real_crush(False)
        R.free_world(w)
    for parent in (pybind.REF_DATA_PARENT, hostlib.DATA_PARENT):  # each side loads it from its own data/
        with open(os.path.join(parent, "data", name), "wb") as f:
            f.write(data)


@needs_ref
def test_free_running_crush_handlers_emulated():
    _crush_lab_sav("_lab_freerun_crush.sav")
    c, n = free_run("emu", "_lab_freerun_crush.sav", (), 0, 40, serialize_every=5)
    assert c["events_crush"] >= 4, c      # all four victims ran their crush handler on the host, mid-doPhysics
    assert c["spawned"] >= 128, c          # two flags exploded into 64 entities each


# ---------------------------------------------------------------- GPU: the named tick counts

@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("case,ticks", [(LOBBY, 1000), (SPACETRASH, 10000), (ARENA, 1000)],
                         ids=["lobby_4players_1000", "spacetrash_spawner_10000", "arena_2players_1000"])
def test_free_running_cuda(case, ticks):
    c, n = free_run("cuda", case[0], case[1], case[2], ticks)
    assert c["pulls"] >= ticks and c["spawned"] > 0
    print("free-running %s: %d ticks bit-exact (hash, full hash on host and device, serialize every 100); %s, %d entities at the end"
          % (case[0], ticks, c, n))


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("level", OTHER_LEVELS)
def test_free_running_every_level_cuda(level):
    free_run("cuda", level + ".sav", (1, 2), 0, 300, serialize_every=100)


@needs_ref
@pytest.mark.gpu
def test_free_running_crush_handlers_cuda():
    _crush_lab_sav("_lab_freerun_crush.sav")
    c, n = free_run("cuda", "_lab_freerun_crush.sav", (), 0, 40, serialize_every=5)
    assert c["events_crush"] >= 4 and c["spawned"] >= 128, c
