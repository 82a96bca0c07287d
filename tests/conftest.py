import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import pybind
    pybind.build(ref=os.path.isdir("/root/reference"))
    return pybind.OracleLib()


@pytest.fixture(scope="session")
def ref_lib():
    from oracle import pybind
    if not pybind.have_ref():
        pytest.skip("oracle/_ref not built (no reference tree on this machine)")
    return pybind.RefLib()


@pytest.fixture(scope="session")
def emu():
    import emubind
    emubind.build_emu()
    return emubind


@pytest.fixture(scope="session")
def bt():
    """The product library on a GPU box; fails loudly if it is missing."""
    import bittoss_b200
    bittoss_b200.lib()
    return bittoss_b200
