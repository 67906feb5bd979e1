#!/usr/bin/env python
"""Development aid (GPU box): first evaluation at which the host-stepped NDT align and the oracle part ways for one random guess.
Both sides print every evaluation (transform, sums) as hex floats when LGS_NDT_EVAL_TRACE is set; this script runs them one after
the other with stderr redirected and compares the lines.\n   usage: diag_ndt_eval_diff.py RES GUESS_INDEX (a guess of test_ndt_parity_over_many_random_guesses) | --case FILE.npz (saved by diag_fuzz.py)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["LGS_NDT_EVAL_TRACE"] = "1"
from lidar_graph_slam_b200 import api  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

if sys.argv[1] == "--case":  # a case saved by tests/diag_fuzz.py
    z = np.load(sys.argv[2])
    td, sd, guess = z["target"], z["source"], z["guess"]
    want_res, eps, it, step, method = float(z["res"]), float(z["eps"]), int(z["it"]), float(z["step"]), int(z["method"])
else:
    want_res, want_k = float(sys.argv[1]), int(sys.argv[2])
    eps, it, step, method = 0.01, 64, 0.1, 2
    z = np.load(os.path.join(ROOT, "tests", "golden", "velodyne_pair.npz"))
    td, sd = O.voxel_grid(z["target"], 0.2)["points"], O.voxel_grid(z["source"], 0.2)["points"]
    rel = z["relative"].astype(np.float64)
    rng = np.random.default_rng(20261018)
    guess = None
    for res in (1.0, 2.0):
        for k in range(24):
            scale = 4.0 if k % 8 == 7 else 1.0
            d = np.eye(4)
            ang = rng.uniform(-1, 1, 3) * np.radians([1.0, 1.0, 4.0]) * scale
            cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
            Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]); Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
            d[:3, :3] = Rx @ Ry @ Rz
            d[:3, 3] = rng.uniform(-1, 1, 3) * np.array([0.6, 0.6, 0.1]) * scale
            if (res, k) == (want_res, want_k):
                guess = (d @ rel).astype(np.float32)
objs = []
for kind in ("stepped", "oracle"):
    n = O.NDT() if kind == "oracle" else api.NormalDistributionsTransform()
    n.setResolution(want_res); n.setTransformationEpsilon(eps); n.setMaximumIterations(it); n.setStepSize(step)
    n.setNeighborhoodSearchMethod(method)
    n.setInputTarget(td); n.setInputSource(sd)
    if kind == "stepped":
        n.profile(1)
    objs.append(n)


def traced(fn, path):
    sys.stderr.flush()
    saved = os.dup(2)
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    os.dup2(fd, 2)
    try:
        fn()
    finally:
        os.dup2(saved, 2)
        os.close(fd)


out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
traced(lambda: objs[0].align(guess), os.path.join(out, "trace_gpu.txt"))
traced(lambda: objs[1].align(guess), os.path.join(out, "trace_oracle.txt"))
A = [l.split() for l in open(os.path.join(out, "trace_gpu.txt")) if l.startswith("EV")]
B = [l.split() for l in open(os.path.join(out, "trace_oracle.txt")) if l.startswith("EV")]
print("evaluations logged: gpu %d (f64 Hessian passes included), oracle %d" % (len(A), len(B)))
A = [a for a in A if a[2] != "2"]
for i, (a, b) in enumerate(zip(A, B)):
    Pa, Pb = a[4:10], b[4:10]
    Ta, Tb = a[11:23], b[11:23]
    ns = 7 if a[2] == "1" else 43
    Sa, Sb = [float.fromhex(v) for v in a[25:25 + ns]], [float.fromhex(v) for v in b[25:25 + ns]]
    same_T = Ta == Tb
    print("   pose", " ".join("%.3e" % (float.fromhex(x) - float.fromhex(y)) for x, y in zip(Pa, Pb)))
    rel = max(abs(x - y) / max(abs(y), 1e-300) for x, y in zip(Sa, Sb))
    print("eval %2d mode %s/%s  T %s  max rel diff of sums %.2e  score %.17g vs %.17g" % (i, a[2], b[2], "same" if same_T else "DIFFERENT", rel, Sa[0], Sb[0]))
    if not same_T:
        for x, y in zip(Ta, Tb):
            if x != y:
                print("    T entry", x, y, float.fromhex(x) - float.fromhex(y))
        break
    if rel > 1e-9:
        for k, (x, y) in enumerate(zip(Sa, Sb)):
            if abs(x - y) > 1e-9 * max(abs(y), 1e-300):
                print("    sum %d: %.17g vs %.17g" % (k, x, y))
        break
