"""The C-ABI shared library: builds for sm_100a, loads, exports every declared symbol,
and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bittoss_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import bittoss_b200
    so = bittoss_b200.build()
    L = ctypes.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libbittoss_b200.so does not export %s" % n


def test_sass_is_sm100a_only():
    import subprocess
    import bittoss_b200
    so = bittoss_b200.build()
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(8|9)\d", out)


def test_no_cpu_fallback_without_gpu():
    import bittoss_b200
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    wa = bittoss_b200.synth_boxstack(64)
    with pytest.raises(bittoss_b200.BtError):
        bittoss_b200.World(wa)


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bittoss_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/ref_harness.cpp", "").replace("oracle/", "ORACLE_DOC/") or \
                    "import oracle" not in txt and "from oracle" not in txt and "libbtoracle" not in txt and \
                    "libbtref" not in txt, f
