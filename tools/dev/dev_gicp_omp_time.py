#!/usr/bin/env python
"""Development aid (GPU box): wall time of pclomp-GICP aligns on the cfg 2 clouds, persistent functor evaluator on / off
(run twice with LGS_NDT_PERSISTENT=1 / 0)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lidar_graph_slam_b200 import api, synth  # noqa: E402

seq, _ = synth.odometry_sequence(n_sweeps=3)
vg = api.VoxelGrid()
vg.setLeafSize(0.25)
clouds = []
for sweep in seq[:2]:
    vg.setInputCloud(synth.drop_invalid(sweep))
    clouds.append(vg.filter(want_membership=False))
g = api.GeneralizedIterativeClosestPoint()
g.setMaxCorrespondenceDistance(2.0)
g.setMaximumIterations(100)
g.setTransformationEpsilon(0.01)
g.setInputTarget(clouds[0])
g.setInputSource(clouds[1])
g.align()
t0 = time.perf_counter()
for _ in range(10):
    g.align()
dt = (time.perf_counter() - t0) / 10
r = g.result
nf = r.evaluations + r.line_search_trials
print("LGS_NDT_PERSISTENT=%s: align %.3f ms, %d outer iterations, %d functor evaluations -> %.1f us per evaluation; T[12:15] = %s" % (
    os.environ.get("LGS_NDT_PERSISTENT", "default"), 1e3 * dt, r.iterations, nf, 1e6 * dt / nf, np.array(r.T)[12:15]))
