#!/usr/bin/env python
"""Development aid (GPU box): per-stage host wall times of loop-closure pairs taken from the key-frame array
(LGS_BATCH_TRACE=1 synchronises after every stage, so the numbers are a breakdown, not the production timeline)."""
import os
import sys

os.environ["LGS_BATCH_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lidar_graph_slam_b200 import api, synth  # noqa: E402

d = synth.loop_keyframes(n_pairs=6, n_keyframes=41, n_azimuth=900, n_unique=2)
kf = api.KeyFrameArray()
for c, P in zip(d["clouds"], d["poses"]):
    kf.push(c, P)
kf.batch_align(d["scan_ids"][:2], d["center_ids"][:2], n_workers=1)
print("---- warm", file=sys.stderr)
kf.batch_align(d["scan_ids"], d["center_ids"], n_workers=1)
