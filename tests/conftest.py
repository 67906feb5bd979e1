import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def velodyne_pair():
    z = np.load(os.path.join(GOLDEN, "velodyne_pair.npz"))
    return dict(target=z["target"], source=z["source"], relative=z["relative"])


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def api():
    """The product API on cuda:0.  Fails (never skips, never falls back) if the extension or the GPU is missing."""
    from lidar_graph_slam_b200 import api as _api
    _api.default_context()
    return _api


def pose_error(T_ref, T):
    """Translation (m) and rotation (rad) between two poses.  The angle comes from the skew part of the relative rotation
    (atan2 of sine and cosine): acos of the trace alone cannot resolve less than sqrt(2 ulp) = 3.4e-4 rad of an f32 matrix."""
    d = np.linalg.inv(np.asarray(T_ref, np.float64)) @ np.asarray(T, np.float64)
    R = d[:3, :3]
    sin_a = 0.5 * np.sqrt((R[2, 1] - R[1, 2]) ** 2 + (R[0, 2] - R[2, 0]) ** 2 + (R[1, 0] - R[0, 1]) ** 2)
    return float(np.linalg.norm(d[:3, 3])), float(np.arctan2(sin_a, (np.trace(R) - 1) / 2))
