#!/usr/bin/env python
"""Development aid (GPU box): wall time of pcl::IterativeClosestPoint aligns (GBS:142-151 settings) on the cfg 2 clouds."""
import os
import sys
import time

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
g = api.IterativeClosestPoint()
g.setMaxCorrespondenceDistance(30)
g.setMaximumIterations(100)
g.setTransformationEpsilon(1e-8)
g.setEuclideanFitnessEpsilon(1e-6)
g.setInputTarget(clouds[0])
g.setInputSource(clouds[1])
g.align()
t0 = time.perf_counter()
for _ in range(10):
    g.align()
dt = (time.perf_counter() - t0) / 10
print("LGS_NDT_PERSISTENT=%s: ICP align %.3f ms, %d iterations -> %.1f us per iteration, state %d" % (
    os.environ.get("LGS_NDT_PERSISTENT", "default"), 1e3 * dt, g.result.iterations, 1e6 * dt / max(g.result.iterations, 1), g.result.line_search_trials))
