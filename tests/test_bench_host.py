"""Host-side pieces of bench.py that can be checked without a GPU: the nvidia-smi clock sampler
(readings before the mark are dropped, throttle reasons are collected, a missing nvidia-smi is
reported instead of raised)."""
import importlib.util
import os
import stat
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_clock_sampler_counts_from_the_mark(tmp_path, monkeypatch):
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nsleep 0.2\nwhile true; do echo '1965, 1965, 640.1, Not Active, Not Active, "
                    "Not Active, Active'; sleep 0.05; done\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    s = _bench().ClockSampler(0)
    s.start()
    time.sleep(0.5)          # "warm-up": nvidia-smi comes up, early readings arrive
    early = len(s.rows)
    s.mark()
    assert early > 0 and s.since_mark() == 0
    time.sleep(0.4)          # the load of interest
    assert s.since_mark() >= 3
    c = s.stop()
    assert c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"]
    assert 3 <= c["samples"] <= len(s.rows) - early


def test_clock_sampler_without_nvidia_smi(tmp_path, monkeypatch):
    monkeypatch.setenv("PATH", str(tmp_path))  # nothing there
    s = _bench().ClockSampler(0)
    s.start()
    s.mark()
    c = s.stop()
    assert c["sm_mhz"] is None and c["reasons"] == ["nvidia-smi unavailable"]
