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
    """nothing under bittoss_b200/, host/ or include/ may import, load, link or execute the oracle; naming its files in
    a comment (where the checker lives) is the only thing allowed"""
    import re
    forbidden = re.compile(r"(import\s+oracle|from\s+oracle|libbtoracle|libbtref|bt_oracle\.c|pybind|dlopen\([^)]*oracle|#include\s+[\"<][^\">]*oracle)")
    checked = 0
    for top in ("bittoss_b200", "host", "include"):
        for dirpath, dirs, files in os.walk(os.path.join(ROOT, top)):
            dirs[:] = [d for d in dirs if d not in ("_build", "__pycache__")]
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c")) and f != "Makefile":
                    continue
                txt = open(os.path.join(dirpath, f)).read()
                checked += 1
                m = forbidden.search(txt)
                assert m is None, "%s/%s names the oracle: %r" % (top, f, m.group(0))
    assert checked >= 10
