#!/usr/bin/env python
"""Development aid (GPU box): NDT target build on the cfg 0 map, for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from lidar_graph_slam_b200 import api, synth  # noqa: E402

d = synth.ndt_scan_to_map()
n = api.NormalDistributionsTransform()
n.setResolution(1.0)
t = torch.from_numpy(d["target"]).cuda()
for _ in range(3):
    n.setInputTarget(t)
torch.cuda.synchronize()
v = n.export_voxels()
print("voxels", len(v["idx"]), "largest", int(v["n"].max()), "mean members", float(v["n"][v["n"] > 0].mean()))
