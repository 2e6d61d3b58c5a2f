"""Runs the C++ host-mirror test (tests/cpp/test_gfs_api.cpp: reference-style test bodies over triqs_b200/gfs.hpp)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cpp", "test_gfs_api")


def test_cpp_header_compiles_cpu_only():
    """the mirror is plain C++17 over the C ABI: it builds on the CPU-only box"""
    assert os.path.exists(EXE), "run __graft_entry__.build()"


@pytest.mark.gpu
def test_cpp_api_on_gpu():
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
    assert "ALL C++ API TESTS PASSED" in r.stdout
