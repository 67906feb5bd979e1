#!/usr/bin/env python
"""Development aid (GPU box): CUDA-event time of the cfg 1 prefilter on a device-resident 262 144-point sweep and of the
NDT target build on the 1 M-point cfg 0 map."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lidar_graph_slam_b200 import api, synth  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = api.Context(0, stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


sw = torch.from_numpy(synth.prefilter_sweeps(n_sweeps=1)[0]).cuda()
vg = api.VoxelGrid(ctx)
vg.setLeafSize(0.2)
vg.setRangeCrop(1.0)
vg.setInputCloud(sw)
print("cfg1 prefilter, device-resident sweep: %.3f ms" % timed(lambda: vg.filter()))
z = np.load(os.path.join(ROOT, "tools", "_cache", "cfg0.npz"))
tgt = torch.from_numpy(z["target"]).cuda()
ndt = api.NormalDistributionsTransform(ctx)
ndt.setResolution(1.0)
print("cfg0 NDT target build, 1 M points: %.3f ms" % timed(lambda: ndt.setInputTarget(tgt), reps=10))
