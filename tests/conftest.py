import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box only)")


@pytest.fixture(scope="session")
def simlib():
    """TEST-ONLY kernel-logic simulation: the CUDA sources compiled as C++ with -DCRGPU_SIM (see cr_common.cuh)."""
    so = os.path.join(ROOT, "tests", "sim", "libcrgpu_sim.so")
    subprocess.run(["make", "-C", os.path.join(ROOT, "comprox_b200", "csrc"), "sim"], check=True, capture_output=True)
    from comprox_b200 import api
    return api.load(so)


@pytest.fixture(scope="session")
def gpulib():
    from comprox_b200 import api
    return api.load()
