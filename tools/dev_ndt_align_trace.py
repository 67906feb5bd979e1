#!/usr/bin/env python
"""Development aid (GPU box): host-side timeline of cfg0 aligns (LGS_NDT_TRACE=1): total wall time per align against the
sum of the device round trips (command sent -> result in the mailbox) of its evaluations."""
import os
import sys

import numpy as np

os.environ["LGS_NDT_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_graph_slam_b200 import api  # noqa: E402

z = np.load(os.path.join(ROOT, "tools", "_cache", "cfg0.npz"))
n = api.NormalDistributionsTransform()
n.setResolution(1.0)
n.setStepSize(0.1)
n.setTransformationEpsilon(0.01)
n.setMaximumIterations(64)
n.setInputTarget(z["target"])
n.setInputSource(z["source"])
for _ in range(6):
    n.align(z["guess"])
