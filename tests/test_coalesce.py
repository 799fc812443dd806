"""Gathering of concurrent single-element calls (csrc/coalesce.h, goldilocks_b200_coalesce): the gate itself is host-only code, so
its grouping, hand-back and statistics are exercised here with a stand-in for the batch launch (tests/coalesce/harness.cpp); the
GPU tier runs the real thing (test_gpu_parity.py::test_single_calls_from_many_threads_are_gathered)."""
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def harness():
    out = os.path.join(ROOT, "build", "coalesce_harness")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", out, os.path.join(HERE, "coalesce", "harness.cpp")])
    return out


def run(harness, threads, calls, window_us, max_batch):
    p = subprocess.run([harness, str(threads), str(calls), str(window_us), str(max_batch)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    return json.loads(p.stdout)


def test_every_call_gets_its_own_result_and_calls_share_batches(harness):
    r = run(harness, 16, 50, 200, 4096)
    assert r["wrong"] == 0 and r["calls"] == 800
    assert r["batches"] < r["calls"] // 4 and 2 <= r["largest"] <= 16
    assert r["served_by_another_thread"] > 0


def test_a_lone_caller_is_served_after_the_window(harness):
    r = run(harness, 1, 20, 100, 4096)
    assert r == {"threads": 1, "calls": 20, "batches": 20, "largest": 1, "wrong": 0, "oversized": 0, "served_by_another_thread": 0,
                 "most_batches_at_a_time": 1}


def test_a_full_gathering_does_not_wait_for_the_window(harness):
    # window of 2 s: 8 calls per thread can only finish in time if reaching max_batch ends the wait
    r = run(harness, 8, 8, 2_000_000, 8)
    assert r["wrong"] == 0 and r["oversized"] == 0 and r["batches"] == 8 and r["largest"] == 8


def test_a_busy_device_makes_the_batches_larger_not_more_numerous(harness):
    # 512 callers, window far shorter than a batch: at most three batches run at a time and the gatherings grow instead
    r = run(harness, 512, 8, 20, 4096)
    assert r["wrong"] == 0 and r["calls"] == 4096
    assert r["most_batches_at_a_time"] <= 3 and r["largest"] >= 64 and r["batches"] <= 200, r


def test_library_exports_the_switch():
    import ctypes as C
    lib = C.CDLL(os.path.join(ROOT, "libgoldilocks_b200", "libgoldilocks_b200.so"))
    lib.goldilocks_b200_coalesce.restype = None
    lib.goldilocks_b200_coalesce_stats.restype = None
    a, b, c = C.c_ulonglong(7), C.c_ulonglong(7), C.c_ulonglong(7)
    lib.goldilocks_b200_coalesce(C.c_uint(0), C.c_uint(0))      # off: no device is touched
    lib.goldilocks_b200_coalesce_stats(C.byref(a), C.byref(b), C.byref(c))
    assert (a.value, b.value, c.value) == (0, 0, 0)
