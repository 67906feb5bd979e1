#!/usr/bin/env python
"""Development aid (GPU box): seeded fuzzing of the CUDA path against the oracle - random clouds (clustered, planar, collinear,
duplicated points, huge and tiny coordinates), random leaf sizes / resolutions / neighbourhoods / guesses.  Prints one line per
disagreement and a summary; exit code 1 if anything disagreed.     usage: diag_fuzz.py [n_cases] [seed]"""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lidar_graph_slam_b200 import api  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from conftest import pose_error  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
bad = []


def cloud(n, kind):
    if kind == "uniform":
        p = rng.uniform(-30, 30, (n, 3)) * np.array([1, 1, 0.1])
    elif kind == "clusters":
        c = rng.uniform(-25, 25, (max(2, n // 200), 3)) * np.array([1, 1, 0.15])
        p = c[rng.integers(0, len(c), n)] + rng.normal(0, 0.4, (n, 3))
    elif kind == "planes":  # walls and a floor: degenerate covariances (eigenvalue clamping, VGC:337-353)
        p = rng.uniform(-20, 20, (n, 3))
        w = rng.integers(0, 3, n)
        p[w == 0, 2] = -1.5
        p[w == 1, 0] = np.round(p[w == 1, 0] / 10) * 10
        p[w == 2, 1] = np.round(p[w == 2, 1] / 10) * 10 + rng.normal(0, 1e-3, (w == 2).sum())
    elif kind == "lines":  # collinear members
        t = rng.uniform(-20, 20, n)
        d = rng.integers(0, 3, n)
        p = np.zeros((n, 3))
        p[np.arange(n), d] = t
        p += np.round(rng.uniform(-3, 3, (n, 3))) * np.array([2.0, 2.0, 0.5]) * (np.arange(3) != d[:, None])
    elif kind == "dupes":
        base = rng.uniform(-10, 10, (max(1, n // 8), 3))
        p = base[rng.integers(0, len(base), n)]
    elif kind == "far":
        p = rng.uniform(-30, 30, (n, 3)) * np.array([1, 1, 0.1]) + np.array([4000.0, -2500.0, 300.0])
    else:
        raise ValueError(kind)
    out = np.zeros((n, 4), np.float32)
    out[:, :3] = p
    out[:, 3] = rng.uniform(0, 255, n)
    return out


KINDS = ["uniform", "clusters", "planes", "lines", "dupes", "far"]


def report(case, what, detail):
    bad.append((case, what))
    print("MISMATCH case %d %s: %s" % (case, what, detail), flush=True)


def fuzz_voxelgrid(case):
    kind = KINDS[rng.integers(0, len(KINDS))]
    n = int(rng.integers(1, 60000))
    pts = cloud(n, kind)
    leaf = float(rng.choice([0.1, 0.2, 0.25, 0.3, 0.5, 0.7, 1.0, 2.0]))
    kw = {}
    if rng.random() < 0.4:
        kw["range_min"] = float(rng.uniform(0.5, 15))
    if rng.random() < 0.3:
        kw["min_pts"] = int(rng.integers(1, 5))
    if rng.random() < 0.3:
        kw["box"] = [-15, 20, -10, 12, -2, 3]
    vg = api.VoxelGrid()
    vg.setLeafSize(leaf)
    if "range_min" in kw:
        vg.setRangeCrop(kw["range_min"])
    if "box" in kw:
        vg.setBoxCrop(kw["box"])
    if "min_pts" in kw:
        vg.setMinimumPointsNumberPerVoxel(kw["min_pts"])
    vg.setInputCloud(pts)
    out = vg.filter()
    ref = O.voxel_grid(pts, leaf, min_points_per_voxel=kw.get("min_pts", 0), range_min=kw.get("range_min", -1.0), box=kw.get("box"))
    tag = "voxelgrid %s n=%d leaf=%g %s" % (kind, n, leaf, kw)
    if vg.info.status != ref["status"] or vg.info.n_kept != ref["n_kept"] or out.shape != ref["points"].shape:
        return report(case, tag, "status/n_kept/shape %s %s %s vs %s %s %s" % (vg.info.status, vg.info.n_kept, out.shape, ref["status"], ref["n_kept"], ref["points"].shape))
    if not np.array_equal(vg.voxel_idx, ref["voxel_idx"]) or not np.array_equal(vg.member_rank, ref["member_rank"]):
        return report(case, tag, "voxel index / membership differ")
    if not np.array_equal(out, ref["points"]):
        return report(case, tag, "centroids differ in %d entries, max %g" % ((out != ref["points"]).sum(), np.abs(out - ref["points"]).max()))


def fuzz_ndt(case, large=False):
    kind = KINDS[rng.integers(0, 5)]  # not "far": the NDT cell table refuses absurd extents by design, tested elsewhere
    nt = int(rng.integers(200, 60000)) if not large else int(rng.integers(160000, 400000))
    tgt = cloud(nt, kind)
    # the source: the target moved by a small rigid motion, subsampled, plus noise
    ang = rng.uniform(-1, 1, 3) * np.radians([1, 1, 3])
    cz, sz = np.cos(ang[2]), np.sin(ang[2])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    ns = int(rng.integers(50, min(nt, 20000) + 1)) if not large else int(rng.integers(155000, nt))  # large: several tiles per CTA
    pick = rng.choice(nt, ns, replace=False)
    src = tgt[pick].copy()
    src[:, :3] = (src[:, :3] - rng.uniform(-0.4, 0.4, 3)) @ Rz + rng.normal(0, 0.02, (ns, 3))
    src = src.astype(np.float32)
    res = float(rng.choice([0.5, 1.0, 1.5, 2.0, 3.0]))
    method = int(rng.choice([1, 2, 3]))
    eps = float(rng.choice([0.01, 0.05, 0.001]))
    it = int(rng.choice([5, 30, 64]))
    step = float(rng.choice([0.1, 0.05, 0.5]))
    g, o = api.NormalDistributionsTransform(), O.NDT()
    for x in (g, o):
        x.setResolution(res)
        x.setTransformationEpsilon(eps)
        x.setMaximumIterations(it)
        x.setStepSize(step)
        x.setNeighborhoodSearchMethod(method)
        x.setInputTarget(tgt)
        x.setInputSource(src)
    tag = "ndt %s nt=%d ns=%d res=%g method=%d eps=%g it=%d step=%g" % (kind, nt, ns, res, method, eps, it, step)
    vg, vo = g.export_voxels(), o.export_voxels()
    if not (np.array_equal(vg["idx"], vo["idx"]) and np.array_equal(vg["n"], vo["n"])):
        return report(case, tag, "occupancy / counts differ")
    valid = vo["n"] >= 6
    if valid.any():
        ig, io = vg["icov"][valid].reshape(valid.sum(), -1), vo["icov"][valid].reshape(valid.sum(), -1)
        e = np.abs(ig - io) / (np.abs(io).max(axis=1, keepdims=True) + 1e-300)
        if e.max() > 1e-9:
            report(case, tag, "icov rel diff %.3e (voxel with n=%d)" % (e.max(), vo["n"][valid][np.unravel_index(e.argmax(), e.shape)[0]]))
    guess = np.eye(4, dtype=np.float32)
    guess[:3, 3] = rng.uniform(-0.3, 0.3, 3) * np.array([1, 1, 0.1])
    for G in (None, guess):
        g.align(G)
        o.align(G)
        r = g.result
        a = (r.iterations, bool(r.converged), r.evaluations, r.line_search_trials, r.hessian_recomputes)
        b = (o.nr_iterations, bool(o.converged), o.stats["derivative_evals"], o.stats["line_search_trials"], o.stats["hessian_recomputes"])
        t_err, r_err = pose_error(o.final_transformation, g.getFinalTransformation())
        bitwise = os.environ.get("LGS_FUZZ_BITWISE") and not np.array_equal(o.final_transformation, g.getFinalTransformation())
        if a != b or not (t_err < 1e-6 and r_err < 1e-6) or bitwise:  # LGS_FUZZ_BITWISE: with LGS_NDT_EXACT_SOLVE=1 the 16 floats are the oracle's
            report(case, tag, "align %s vs %s  pose diff %.3e m %.3e rad" % (a, b, t_err, r_err))
            out = os.path.join(ROOT, "gpurun_out")
            if os.path.isdir(out):  # input of tests/diag_ndt_eval_diff.py --case
                np.savez_compressed(os.path.join(out, "fuzz_ndt_case_%d_%d.npz" % (seed, case)), target=tgt, source=src, res=res, method=method, eps=eps,
                                    it=it, step=step, guess=np.eye(4, dtype=np.float32) if G is None else G)
        else:
            fg, fo = g.getFitnessScore(), o.getFitnessScore()
            if not abs(fg - fo) <= 1e-9 * abs(fo):
                report(case, tag, "fitness %.17g vs %.17g" % (fg, fo))


def fuzz_gicp(case):
    kind = KINDS[rng.integers(0, 4)]
    nt = int(rng.integers(100, 30000))
    tgt = cloud(nt, kind)
    ns = int(rng.integers(50, min(nt, 10000) + 1))
    pick = rng.choice(nt, ns, replace=False)
    src = tgt[pick].copy()
    yaw = rng.uniform(-1, 1) * np.radians(3)
    cz, sz = np.cos(yaw), np.sin(yaw)
    src[:, :3] = (src[:, :3] - rng.uniform(-0.3, 0.3, 3)) @ np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) + rng.normal(0, 0.02, (ns, 3))
    src = src.astype(np.float32)
    k = int(rng.choice([10, 20]))
    corr = float(rng.choice([1.0, 2.0, 1e30]))
    reg = int(rng.integers(0, 5))
    g, o = api.FastGICP(), O.FastGICP()
    for x in (g, o):
        x.setCorrespondenceRandomness(k)
        x.setMaxCorrespondenceDistance(corr)
        x.setRegularizationMethod(reg)
        x.setMaximumIterations(64)
        x.setTransformationEpsilon(0.01)
        x.setInputTarget(tgt)
        x.setInputSource(src)
    tag = "gicp %s nt=%d ns=%d k=%d corr=%g reg=%d" % (kind, nt, ns, k, corr, reg)
    for which in (0, 1):
        cg, co = g.covariances(which), o.covariances(which)
        cg, co = cg.reshape(len(cg), -1), co.reshape(len(co), -1)
        sc = np.abs(co).max(axis=1, keepdims=True) + 1e-300
        e = (np.abs(cg - co) / sc).max()
        if not e <= 1e-9:
            report(case, tag, "covariances[%d] rel diff %.3e" % (which, e))
    g.align()
    o.align()
    a = (g.result.iterations, bool(g.result.converged), g.result.evaluations, g.result.line_search_trials)
    b = (o.nr_iterations, bool(o.converged), o.stats["linearize_calls"], o.stats["error_calls"])
    t_err, r_err = pose_error(o.final_transformation, g.getFinalTransformation())
    if a != b or not (t_err < 1e-4 and r_err < 1e-4):
        report(case, tag, "align %s vs %s  pose diff %.3e m %.3e rad" % (a, b, t_err, r_err))


def fuzz_ndt_large(case):
    if case % 20 == 7:
        fuzz_ndt(case, large=True)


def _pair_for_icp():
    kind = KINDS[rng.integers(0, 3)]
    nt = int(rng.integers(300, 20000))
    tgt = cloud(nt, kind)
    ns = int(rng.integers(100, min(nt, 8000) + 1))
    src = tgt[rng.choice(nt, ns, replace=False)].copy()
    yaw = rng.uniform(-1, 1) * np.radians(3)
    cz, sz = np.cos(yaw), np.sin(yaw)
    src[:, :3] = (src[:, :3] - rng.uniform(-0.3, 0.3, 3)) @ np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) + rng.normal(0, 0.02, (ns, 3))
    return kind, tgt, src.astype(np.float32)


def fuzz_icp(case):
    if case % 3:
        return
    kind, tgt, src = _pair_for_icp()
    corr = float(rng.choice([1.0, 2.0, 30.0]))
    g, o = api.IterativeClosestPoint(), O.IterativeClosestPoint()
    for x in (g, o):
        x.setMaxCorrespondenceDistance(corr)
        x.setMaximumIterations(100)
        x.setTransformationEpsilon(1e-8)
        x.setEuclideanFitnessEpsilon(1e-6)
        x.setInputTarget(tgt)
        x.setInputSource(src)
    tag = "icp %s nt=%d ns=%d corr=%g" % (kind, len(tgt), len(src), corr)
    for _ in range(2):  # twice: PCL's convergence criteria keep state between aligns
        g.align()
        o.align()
        a = (g.result.iterations, bool(g.result.converged), g.result.line_search_trials)
        b = (o.nr_iterations, bool(o.converged), o.stats["convergence_state"])
        t_err, r_err = pose_error(o.final_transformation, g.getFinalTransformation())
        if a != b or not (t_err < 1e-4 and r_err < 1e-4):
            report(case, tag, "align %s vs %s  pose diff %.3e m %.3e rad" % (a, b, t_err, r_err))


def fuzz_gicp_omp(case):
    if case % 3 != 1:
        return
    kind, tgt, src = _pair_for_icp()
    corr = float(rng.choice([1.0, 2.0]))
    g, o = api.GeneralizedIterativeClosestPoint(), O.GeneralizedIterativeClosestPoint()
    for x in (g, o):
        x.setMaxCorrespondenceDistance(corr)
        x.setTransformationEpsilon(0.01)
        x.setMaximumIterations(30)
        x.setMaximumOptimizerIterations(10)
        x.setInputTarget(tgt)
        x.setInputSource(src)
    tag = "gicp_omp %s nt=%d ns=%d corr=%g" % (kind, len(tgt), len(src), corr)
    g.align()
    o.align()
    a = (g.result.iterations, bool(g.result.converged), g.result.line_search_trials, g.result.evaluations, g.result.hessian_recomputes)
    b = (o.nr_iterations, bool(o.converged), o.stats["f_calls"], o.stats["df_calls"] + o.stats["fdf_calls"], o.stats["inner_iterations"])
    t_err, r_err = pose_error(o.final_transformation, g.getFinalTransformation())
    if a != b or not (t_err < 1e-4 and r_err < 1e-4):
        report(case, tag, "align %s vs %s  pose diff %.3e m %.3e rad" % (a, b, t_err, r_err))


FUZZERS = (fuzz_voxelgrid, fuzz_ndt, fuzz_gicp)
if os.environ.get("LGS_FUZZ_MORE"):  # the slower fuzzers: multi-tile NDT sources, ICP, pclomp GICP
    FUZZERS = FUZZERS + (fuzz_ndt_large, fuzz_icp, fuzz_gicp_omp)
for case in range(n_cases):
    for fn in FUZZERS:
        try:
            fn(case)
        except Exception as e:  # a refusal on one side only is a disagreement too
            report(case, fn.__name__, "exception %r\n%s" % (e, traceback.format_exc(limit=3)))
print("%d cases x %d fuzzers, %d disagreements" % (n_cases, len(FUZZERS), len(bad)))
sys.exit(1 if bad else 0)
