"""Runs the C++ tests of the facade (include/ttessel.h): tests/cpp/test_facade.cpp,
written after the reference's own Boost.Test suite (reference tests/check_lite.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_facade")


def _build():
    subprocess.check_call(["make", "-C", ROOT, "tests/cpp/test_facade"], stdout=subprocess.DEVNULL)


def test_facade_host_part():
    """Split known answer, the reference's `squared_domain` face-prediction property
    (tests/check_lite.cpp:1031-1040), text round trip of the 11-cell fixture, CatVector
    algebra, feature registry errors, host feature functions."""
    _build()
    r = subprocess.run([EXE, "cpu", os.path.join(ROOT, "tests", "golden", "tessellation_data.txt")],
                       capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_facade_device_part():
    """Energy::statistic_variation on the GPU against the host element lists for all 12
    features; CRTT Newton fit to the closed form; ACS value from AddGivenSplit."""
    if not os.path.exists(EXE):
        _build()
    r = subprocess.run([EXE, "gpu"], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
