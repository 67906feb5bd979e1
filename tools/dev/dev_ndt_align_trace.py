#!/usr/bin/env python
"""Development aid (GPU box): where a device-resident NDT align spends its time (CTA 0's clock64 breakdown) on cfg 0."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from lidar_graph_slam_b200 import api, synth  # noqa: E402

d = synth.ndt_scan_to_map()
n = api.NormalDistributionsTransform()
n.setResolution(1.0)
n.setStepSize(0.1)
n.setTransformationEpsilon(0.01)
n.setMaximumIterations(64)
n.setInputTarget(d["target"])
n.setInputSource(torch.from_numpy(d["source"]).cuda())
for rep in range(3):
    t0 = time.perf_counter()
    n.align(d["guess"])
    dt = time.perf_counter() - t0
    r, b = n.result, n.align_breakdown()
    ev = r.evaluations + r.hessian_recomputes
    print("align %.1f us wall: %d iterations, %d evaluations; per evaluation (cycles): evaluate %.0f  wait %.0f  add rows %.0f  optimiser %.0f  publish %.0f; kernel %.0f cycles = %.1f us at 1.965 GHz"
          " | optimiser: decide %.0f solve %.0f after-solve %.0f trig %.0f pose+tables %.0f"
          % (1e6 * dt, r.iterations, ev, b["evaluate"] / ev, b["wait_grid"] / ev, b["add_rows"] / ev, b["optimiser"] / ev, b["publish"] / ev, b["total"], b["total"] / 1965.0,
             b["opt_decide"] / ev, b["opt_solve"] / ev, b["opt_after_solve"] / ev, b["opt_trig"] / ev, b["opt_pose_tables"] / ev))
    raw = b["raw"]
    print("   CTA 0 per evaluation: body %.0f  idle (arrival -> next command) %.0f  command load %.0f" % (raw[0] / ev, raw[6] / ev, raw[7] / ev))
