#!/usr/bin/env python
"""Development aid (GPU box): iteration-count parity of the NDT align with the oracle over random guesses, for the device-resident
align and the host-stepped one (LGS_NDT_FORCE_SVD=1 makes the host-stepped Newton step a JacobiSVD like the reference's)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_graph_slam_b200 import api  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "velodyne_pair.npz"))
td, sd = O.voxel_grid(z["target"], 0.2)["points"], O.voxel_grid(z["source"], 0.2)["points"]
rel = z["relative"].astype(np.float64)
rng = np.random.default_rng(20261018)
for res in (1.0, 2.0):
    objs = []
    for kind in ("device", "stepped", "oracle"):
        n = O.NDT() if kind == "oracle" else api.NormalDistributionsTransform()
        n.setResolution(res); n.setTransformationEpsilon(0.01); n.setMaximumIterations(64); n.setStepSize(0.1)
        n.setInputTarget(td); n.setInputSource(sd)
        if kind == "stepped":
            n.profile(1)
        objs.append(n)
    dev, stp, orc = objs
    bad_dev = bad_stp = 0
    for k in range(24):
        scale = 4.0 if k % 8 == 7 else 1.0
        d = np.eye(4)
        ang = rng.uniform(-1, 1, 3) * np.radians([1.0, 1.0, 4.0]) * scale
        cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]); Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        d[:3, :3] = Rx @ Ry @ Rz
        d[:3, 3] = rng.uniform(-1, 1, 3) * np.array([0.6, 0.6, 0.1]) * scale
        guess = (d @ rel).astype(np.float32)
        dev.align(guess); stp.align(guess); orc.align(guess)
        a = (dev.result.iterations, dev.result.evaluations, dev.result.line_search_trials, dev.result.hessian_recomputes)
        b = (stp.result.iterations, stp.result.evaluations, stp.result.line_search_trials, stp.result.hessian_recomputes)
        c = (orc.nr_iterations, orc.stats["derivative_evals"], orc.stats["line_search_trials"], orc.stats["hessian_recomputes"])
        dT = np.abs(dev.getFinalTransformation() - orc.final_transformation).max()
        sT = np.abs(stp.getFinalTransformation() - orc.final_transformation).max()
        flag = ("" if a == c else " DEVICE!=ORACLE") + ("" if b == c else " STEPPED!=ORACLE")
        bad_dev += a != c; bad_stp += b != c
        print("res %.1f guess %2d: device %s stepped %s oracle %s  max|dT| %.2e / %.2e%s" % (res, k, a, b, c, dT, sT, flag))
    print("res %.1f: %d / 24 device mismatches, %d / 24 stepped mismatches (LGS_NDT_FORCE_SVD=%s)" % (res, bad_dev, bad_stp, os.environ.get("LGS_NDT_FORCE_SVD")))
