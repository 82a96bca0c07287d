"""bittoss_b200.savfile (direct .sav -> SoA) against the reference's own deserialize + export,
on every shipped level."""
import glob
import os

import numpy as np
import pytest

from oracle import pybind

LEVELS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(pybind.REF_DATA_PARENT, "data", "*")) if os.path.isfile(p))


@pytest.mark.parametrize("name", LEVELS or ["lobby.sav"])
def test_sav_reader_matches_reference_loader(ref_lib, name):
    from bittoss_b200.savfile import load_sav
    path = os.path.join(pybind.REF_DATA_PARENT, "data", name)
    wa, info = load_sav(path)
    R = ref_lib
    wid = R.new_world()
    R.load_sav(name)
    ra = R.export()
    rules = R.get_gamerules()
    R.free_world(wid)
    assert wa.n == ra.n and info["gamerules"] == rules
    a, b = wa.state_dict(), ra.state_dict()
    for k in a:
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k
    for k in ("root_order", "holdee_head", "holdee_next"):
        assert np.array_equal(wa.a[k][: wa.n], ra.a[k][: ra.n]), k


def test_sav_reader_rejects_garbage():
    from bittoss_b200.savfile import SavError, load_sav
    with pytest.raises(SavError):
        load_sav(b"not a save file at all")
    with pytest.raises(SavError):
        load_sav(b"gam7" + b"\x00\x00\x10\x00" + b"\x00" * 20)


@pytest.mark.parametrize("name", LEVELS or ["lobby.sav"])
def test_sav_writer_matches_reference_writer(ref_lib, name):
    """save_sav(export of the reference's state) == the reference's own serialize(), byte for byte:
    right after the load, and again after the level was kicked and stepped (moved entities, changed
    sliders, deaths, new hold trees)."""
    from bittoss_b200.savfile import load_sav, save_sav
    R = ref_lib
    wid = R.new_world()
    R.load_sav(name)
    try:
        for phase in range(2):
            if phase:
                R.strip_handlers(drop_ticks=True)
                R.kick(5, 300, 3)
                for _ in range(25):
                    R.step(1)
            want = R.serialize()
            wa = R.export()
            _, info = load_sav(want)  # colours / players / rand are pass-through fields
            got = save_sav(wa, colors=info["colors"], players=info["players"], rand=info["rand"],
                           gamerules=R.get_gamerules(), gamestrings=info["gamestrings"])
            assert got == want, "phase %d: first difference at byte %d of %d/%d" % (
                phase, next((i for i, (x, y) in enumerate(zip(got, want)) if x != y), min(len(got), len(want))),
                len(got), len(want))
    finally:
        R.free_world(wid)


@pytest.mark.parametrize("name", LEVELS or ["lobby.sav"])
def test_sav_write_read_round_trip(name):
    """load -> save -> load gives the same arrays, whatever version the file on disk has."""
    from bittoss_b200.savfile import load_sav, save_sav
    path = os.path.join(pybind.REF_DATA_PARENT, "data", name)
    if not os.path.isfile(path):
        pytest.skip("no shipped levels on this machine")
    wa, info = load_sav(path)
    data = save_sav(wa, colors=info["colors"], players=info["players"], rand=info["rand"],
                    gamestrings=info["gamestrings"])
    wb, info_b = load_sav(data)
    assert info_b["version"] == 7 and info_b["colors"] == info["colors"] and info_b["gamerules"] == info["gamerules"]
    assert info_b["gamestrings"] == [g for g in info["gamestrings"]][: len(info_b["gamestrings"])]
    a, b = wa.state_dict(), wb.state_dict()
    for k in a:
        if k == "btn":
            continue  # a file carries one button byte: it loads as "pressed next tick" (v2), is written from v
        assert np.array_equal(a[k], b[k]), k


def test_sav_writer_refuses_a_batch_and_keeps_passthrough_fields():
    """the format has no world id: a batched desc must be split first; colours, players, rand and
    gamestrings handed to the writer come back from the reader unchanged"""
    from bittoss_b200.capi import concat_worlds
    from bittoss_b200.desc import WorldArrays
    from bittoss_b200.savfile import SavError, load_sav, save_sav
    one = WorldArrays(2, 1, 1, 1)
    a = one.a
    a["uid"][:2] = (0, 1)
    a["center"][:6] = (1, 2, 3, -4, -5, -6)
    a["radius"][:6] = 100
    a["holder"][:2] = -1
    a["holdee_head"][:2] = -1
    a["holdee_next"][:2] = -1
    a["new_holder"][:2] = -1
    a["friction"][:2] = (16, 3)
    a["slider_off"][:3] = (0, 1, 1)
    a["slider_v"][0] = 77
    a["wire_off"][:3] = (0, 0, 1)
    a["wire_to"][0] = 0
    a["handlers"][10 + 4] = 5  # entity 1: push handler id 5 (file position 4)
    players = [{"color": 0x112233, "revive": 9, "data": 4, "name": b"ann", "entity": 1}]
    data = save_sav(one, colors=[7, -8], players=players, rand=424242, gamerules=12, gamestrings=[b"hello", b"", b"x", b""])
    wb, info = load_sav(data)
    assert info["colors"] == [7, -8] and info["players"] == players and info["rand"] == 424242
    assert info["gamerules"] == 12 and info["gamestrings"] == [b"hello", b"", b"x"]
    assert list(wb.a["center"][:6]) == [1, 2, 3, -4, -5, -6] and list(wb.a["friction"][:2]) == [16, 3]
    assert wb.a["slider_v"][0] == 77 and wb.a["wire_to"][0] == 0 and wb.a["handlers"][10 + 4] == 5
    with pytest.raises(SavError):
        save_sav(concat_worlds([one, one]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in LEVELS if n.endswith(".sav")] or ["lobby.sav"])
def test_sav_device_round_trip_matches_reference(ref_lib, name):
    """SURVEY §8(f)-3 end to end: .sav -> SoA (load_sav, no `ent` in between) -> kicked and stepped ON THE DEVICE ->
    download -> save_sav == the bytes the reference's own serialize() writes after the same kicks and ticks on the CPU."""
    import bittoss_b200
    from bittoss_b200.savfile import load_sav, save_sav
    path = os.path.join(pybind.REF_DATA_PARENT, "data", name)
    wa, info = load_sav(path)
    wa.a["handlers"][1::10] = 0          # tick / tickHeld stay out of it on both sides (physics-only steps)
    wa.a["handlers"][2::10] = 0
    R = ref_lib
    wid = R.new_world()
    R.load_sav(name)
    try:
        R.strip_handlers(drop_ticks=True, keep_push=0x3ff, keep_pushed=0x1ff, keep_pair=0x7f, keep_crush=0)
        R.kick(5, 300, 3)
        kicked = R.export()               # the kicks are pending d_vel: carry exactly those over
        wa.a["d_vel"][: 3 * wa.n] = kicked.a["d_vel"][: 3 * wa.n]
        wa.a["handlers"][3::10] = 0       # crush handlers: explode() would create entities on the host
        w = bittoss_b200.World(wa)
        for _ in range(25):
            R.step(1)           # physics only on both sides (bt_world_tick would add the world rules of the level)
            w.step(31)
        assert w.hash()[0] == R.hash()
        out = w.download()
        w.close()
        want = R.serialize()
        _, winfo = load_sav(want)         # colours / players / rand are pass-through fields of the live entities
        got = save_sav(out, colors=winfo["colors"], players=winfo["players"], rand=winfo["rand"],
                       gamerules=R.get_gamerules(), gamestrings=winfo["gamestrings"])
        # the reference names its handlers by catalog id: with tick handlers dropped on both sides the files agree
        assert got == want, "first difference at byte %d of %d/%d" % (
            next((i for i, (x, y) in enumerate(zip(got, want)) if x != y), min(len(got), len(want))), len(got), len(want))
    finally:
        R.free_world(wid)
