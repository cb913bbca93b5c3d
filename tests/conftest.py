import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def hs():
    """The product binding (ctypes over libhsgpu.so). Building is nvcc-only, no GPU needed."""
    import tfe_slam_ucl_b200 as m
    if not os.path.exists(m.LIB_PATH):
        m.build()
    m.lib()
    return m


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    if not pyoracle.have("port"):
        pyoracle.build(ref=os.path.isdir("/root/reference"))
    return pyoracle


@pytest.fixture(scope="session")
def have_ref(oracle):
    return oracle.have("ref")
