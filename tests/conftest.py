"""Shared fixtures.  `-m "not gpu"` runs on a CPU-only box (oracle, host logic, ABI surface);
`-m gpu` are the parity tests proper: CUDA path vs oracle, through the C ABI, on a B200."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def b200slam():
    mod = importlib.import_module("hardware-acceleration-of-lidar-slam_b200")
    if not os.path.exists(mod.LIB_PATH):
        mod.build()
    return mod


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("hardware-acceleration-of-lidar-slam_b200.synth")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    if not os.path.exists(os.path.join(pyoracle.HERE, "liboracle.so")):
        pyoracle.build()
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ctx(b200slam):
    """GPU context.  No skip: on a GPU box a failure to create it must fail the test."""
    c = b200slam.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def edt_golden():
    return np.load(os.path.join(GOLDEN, "edt_golden.npz"))


@pytest.fixture(scope="session")
def fastmatch_golden():
    return np.load(os.path.join(GOLDEN, "fastmatch_golden.npz"))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(params=["fastmatch_kernel", "lattice_kernel"])
def small_lattice_kernel(request, monkeypatch):
    """FastMatch-sized lattices (<= 32 candidates) run in the one-CTA fastmatch_kernel; B200SLAM_NO_FASTMATCH_KERNEL
    (read at every launch) sends them through the general lattice kernel's per-candidate-count instantiation, which
    larger small lattices and long scans still use.  Tests of the FastMatch semantics run under both."""
    if request.param == "lattice_kernel":
        monkeypatch.setenv("B200SLAM_NO_FASTMATCH_KERNEL", "1")
    else:
        monkeypatch.delenv("B200SLAM_NO_FASTMATCH_KERNEL", raising=False)
    return request.param
