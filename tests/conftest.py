import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def hostsim():
    """ctypes handle of the host-compiled kernel arithmetic (tests/hostsim; test infrastructure only)."""
    import ctypes

    d = os.path.join(ROOT, "tests", "hostsim")
    so, src = os.path.join(d, "_hostsim.so"), os.path.join(d, "hostsim.cpp")
    hdrs = [os.path.join(ROOT, "abr_control_b200", "csrc", f) for f in
            ("abrb_math.cuh", "abrb_rbd.cuh", "abrb_osc.cuh", "abrb_host.hpp")]
    newest = max(os.path.getmtime(p) for p in [src] + hdrs)
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so], check=True)
    return ctypes.CDLL(so)
