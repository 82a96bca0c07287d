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
