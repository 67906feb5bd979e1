#!/usr/bin/env python
"""Diagnostic (test infrastructure: compares against the oracle, hence lives under tests/).  Development loop for the NDT evaluation kernels on the cfg0 workload (run on the GPU box):
parity of one evaluation per mode against the oracle, per-kernel CUDA-event timings, then a full align parity check.

    python tests/diag_ndt_deriv.py [--no-oracle] [--reps 50]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def workload():
    cache = os.path.join(ROOT, "tools", "_cache", "cfg0.npz")
    if os.path.exists(cache):
        z = np.load(cache)
        return {k: z[k] for k in z.files}
    from lidar_graph_slam_b200 import synth
    return synth.ndt_scan_to_map()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--method", type=int, default=None)
    args = ap.parse_args()
    from lidar_graph_slam_b200 import api
    d = workload()
    g = api.NormalDistributionsTransform()
    objs = [g]
    if not args.no_oracle:
        from oracle import pyoracle as O
        o = O.NDT()
        objs.append(o)
    for n in objs:
        n.setResolution(1.0)
        n.setStepSize(0.1)
        n.setTransformationEpsilon(0.01)
        n.setMaximumIterations(64)
        if args.method is not None:
            n.setNeighborhoodSearchMethod(args.method)
        n.setInputTarget(d["target"])
        n.setInputSource(d["source"])
    # evaluation pose = the workload's initial guess, as (x, y, z, roll, pitch, yaw) of T = Trans * Rx * Ry * Rz (NDT.h:214-231)
    G = d["guess"].astype(np.float64)
    ry = np.arcsin(G[0, 2])
    p = np.array([G[0, 3], G[1, 3], G[2, 3], np.arctan2(-G[1, 2], G[2, 2]), ry, np.arctan2(-G[0, 1], G[0, 0])])
    if not args.no_oracle:
        T = O.ndt_convert_transform(p)
        for mode in (0, 1, 2):
            so, go, Ho = o.derivatives(T, p, mode)
            sg, gg, Hg = g.derivatives(T, p, mode)
            scale = np.abs(Ho).max() if mode != 1 else 1.0
            if mode != 2:
                print("mode %d score rel err %.3e  grad max rel err %.3e" % (mode, abs(sg - so) / abs(so), np.abs(gg - go).max() / np.abs(go).max()))
            if mode != 1:
                print("mode %d triu(H) max err / scale %.3e   full H %.3e" % (mode, np.abs(np.triu(Hg) - np.triu(Ho)).max() / scale, np.abs(Hg - Ho).max() / scale))
    else:
        T = d["guess"].astype(np.float32)
    # timings
    g.profile(True)
    for mode in (0, 1, 2):
        for _ in range(args.reps):
            g.derivatives(T, p, mode)
    pr = g.profile(False)
    print("kernel us/launch: hess %.2f  grad %.2f  h64 %.2f   terms %.0f (h_bar %.3f)" % (
        1e3 * pr["hess_ms"] / max(pr["hess_launches"], 1), 1e3 * pr["grad_ms"] / max(pr["grad_launches"], 1),
        1e3 * pr["h64_ms"] / max(pr["h64_launches"], 1), pr["terms_last_eval"], pr["terms_last_eval"] / pr["n_source"]))
    # per-CTA phase stamps (only in a -DLGS_DERIV_TRACE build)
    import ctypes as C
    L = g._L
    if hasattr(L, "lgs_ndt_debug_trace"):
        for mode in (0, 1, 2):
            g.derivatives(T, p, mode)
            tr = np.zeros((148, 8))
            L.lgs_ndt_debug_trace(g._h, tr.ctypes.data_as(C.c_void_p), 148)
            names = ["phase1-loads", "probes+tables", "compaction", "phase2", "scalar", "reduce"]
            dd = np.diff(np.concatenate([np.zeros((148, 1)), tr[:, 1:7]], axis=1), axis=1)
            print("mode %d trace (cycles, mean / max over CTAs): " % mode + "  ".join("%s %.0f/%.0f" % (nm, dd[:, i].mean(), dd[:, i].max()) for i, nm in enumerate(names)))
            print("   total cycles mean %.0f max %.0f ; wall (globaltimer) first start -> last end %.2f us ; start spread %.2f us ; per-CTA duration mean %.2f us" % (
                tr[:, 6].mean(), tr[:, 6].max(), (tr[:, 7].max() - tr[:, 0].min()) / 1e3, (tr[:, 0].max() - tr[:, 0].min()) / 1e3, (tr[:, 7] - tr[:, 0]).mean() / 1e3))
    # launch + event floor: a 32-point source
    t0 = time.perf_counter()
    for _ in range(args.reps):
        g.derivatives(T, p, 0)
    print("host wall per mode-0 evaluation: %.1f us" % (1e6 * (time.perf_counter() - t0) / args.reps))
    # align
    t0 = time.perf_counter()
    for _ in range(10):
        g.align(d["guess"])
    dt = (time.perf_counter() - t0) / 10
    r = g.result
    print("align: %.3f ms  iterations %d evals %d trials %d hess_recomputes %d converged %d" % (1e3 * dt, r.iterations, r.evaluations, r.line_search_trials,
                                                                                              r.hessian_recomputes, r.converged))
    if not args.no_oracle:
        o.align(d["guess"])
        E = np.linalg.inv(o.final_transformation.astype(np.float64)) @ g.getFinalTransformation().astype(np.float64)
        print("oracle: iterations %d evals %d trials %d hess %d ; pose diff %.3e m %.3e rad ; trans_prob rel %.3e" % (
            o.nr_iterations, o.stats["derivative_evals"], o.stats["line_search_trials"], o.stats["hessian_recomputes"],
            np.linalg.norm(E[:3, 3]), np.arccos(np.clip((np.trace(E[:3, :3]) - 1) / 2, -1, 1)),
            abs(g.getTransformationProbability() - o.trans_probability) / abs(o.trans_probability)))


if __name__ == "__main__":
    main()
