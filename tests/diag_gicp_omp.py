#!/usr/bin/env python
"""Diagnostic (test infrastructure: compares against the oracle, hence lives under tests/).  pclomp-GICP parity diagnostics on the synthetic loop pairs: covariances, one functor
set-up, and the BFGS call counts of a whole align, GPU vs oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_graph_slam_b200 import api, synth  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

scans, submaps, corrections = synth.loop_pairs(n_pairs=3, n_keyframes=9, n_unique=2)
for i in range(3):
    tgt = O.voxel_grid(submaps[i], 0.5)["points"]
    g, o = api.GeneralizedIterativeClosestPoint(), O.GeneralizedIterativeClosestPoint()
    for x in (g, o):
        x.setMaxCorrespondenceDistance(2.0)
        x.setMaximumIterations(100)
        x.setMaximumOptimizerIterations(20)
        x.setTransformationEpsilon(0.01)
        x.setInputTarget(tgt)
        x.setInputSource(scans[i])
    for w in (0, 1):
        cg, co = g.covariances(w), o.covariances(w)
        d = np.abs(cg - co).reshape(len(cg), -1).max(1)
        print("pair %d cov[%d] max abs diff %.3e  (rows > 1e-9: %d of %d)" % (i, w, d.max(), (d > 1e-9).sum(), len(d)))
    I = np.eye(4, dtype=np.float32)
    x = np.zeros(6)
    rg, ro = g.functor(I, I, x), o.functor(I, I, x)
    v = ro["corr"] >= 0
    print("   corr equal %s  n_corr %d/%d  mahal max diff %.3e  f rel %.3e  fdf_f rel %.3e  g max rel %.3e" % (
        np.array_equal(rg["corr"], ro["corr"]), rg["n_corr"], ro["n_corr"], np.abs(rg["mahal"][v] - ro["mahal"][v]).max(),
        abs(rg["f"] - ro["f"]) / abs(ro["f"]), abs(rg["fdf_f"] - ro["fdf_f"]) / abs(ro["fdf_f"]),
        np.abs(rg["fdf_g"] - ro["fdf_g"]).max() / np.abs(ro["fdf_g"]).max()))
    o.align()
    g.align()
    print("   oracle: iters %d conv %d %s" % (o.nr_iterations, o.converged, o.stats))
    print("   gpu   : iters %d conv %d f %d df+fdf %d inner %d" % (g.result.iterations, g.result.converged, g.result.line_search_trials,
                                                                    g.result.evaluations, g.result.hessian_recomputes))
    d = np.linalg.inv(o.final_transformation.astype(np.float64)) @ g.getFinalTransformation().astype(np.float64)
    print("   pose diff %.3e m" % np.linalg.norm(d[:3, 3]))
