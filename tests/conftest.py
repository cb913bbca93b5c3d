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


def make_streams(hs, n_streams, n_scans, preset=0, seed0=1000, scale_to_map=20.0, first_scan=0, max_points=None):
    """Synthetic scan streams -> (scene, ranges[k][s][b], points[k][s][mp][2], counts[k][s], truth)."""
    sc = hs.scene_preset(preset)
    ranges, truth = hs.scangen_batch(sc, seed0, n_streams, first_scan, n_scans)
    pts, cnt = hs.scan_to_points_host(ranges, sc, scale_to_map, max_points)
    return sc, ranges, pts, cnt, truth


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)
