#!/usr/bin/env python
"""Development aid (GPU box, under ncu): one GICP covariance pass (self k-NN, k = 20) on a 0.25 m-filtered 64-beam sweep."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lidar_graph_slam_b200 import api, synth  # noqa: E402

seq, _ = synth.odometry_sequence(n_sweeps=3)
vg = api.VoxelGrid()
vg.setLeafSize(0.25)
vg.setInputCloud(synth.drop_invalid(seq[0]))
cloud = vg.filter(want_membership=False)
g = api.FastGICP()
g.setInputTarget(cloud)
c = g.covariances(1)
g.setInputTarget(cloud)
c = g.covariances(1)
print(cloud.shape, c.shape)
