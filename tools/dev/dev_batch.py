#!/usr/bin/env python
"""Development loop for the loop-closure batch (cfg 4 in miniature, run on the GPU box): wall time per pair against
the number of host workers, and the per-stage split of one pair.

    python tools/dev/dev_batch.py [--pairs 8] [--workers 1,2,4,8]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=8)
    ap.add_argument("--workers", default="1,2,4,8")
    ap.add_argument("--method", default="gicp")
    args = ap.parse_args()
    from lidar_graph_slam_b200 import api, synth
    scans, submaps, _ = synth.loop_pairs(n_pairs=args.pairs, n_keyframes=41, n_azimuth=900, n_unique=2)
    print("pairs %d  scan pts %.0f  submap pts %.0f" % (args.pairs, np.mean([len(s) for s in scans]), np.mean([len(s) for s in submaps])))
    method = api.METHOD_GICP if args.method == "gicp" else api.METHOD_NDT
    api.batch_align(scans[:1], submaps[:1], method=method, n_workers=1)
    for w in [int(x) for x in args.workers.split(",")]:
        api.batch_align(scans[:2 * w], submaps[:2 * w], method=method, n_workers=w)  # new workers allocate their device state once
        t0 = time.perf_counter()
        recs = api.batch_align(scans, submaps, method=method, n_workers=w)
        dt = time.perf_counter() - t0
        print("workers %2d: %.1f ms/pair  (%.1f pairs/s)  converged %d  iterations %s" % (
            w, 1e3 * dt / args.pairs, args.pairs / dt, sum(r.converged for r in recs), [r.iterations for r in recs][:8]))
    # stage split of one pair through the single-object API
    import torch
    vg = api.VoxelGrid()
    vg.setLeafSize(0.5)
    g = api.FastGICP()
    g.setMaxCorrespondenceDistance(2.0)

    def lap(label, fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        torch.cuda.synchronize()
        print("  %-34s %.2f ms" % (label, 1e3 * (time.perf_counter() - t0) / reps))
        return out

    def filt():
        vg.setInputCloud(submaps[0])
        return vg.filter()

    ds = lap("VoxelGrid 0.5 m (host in/out)", filt)
    print("  submap after filter: %d pts" % len(ds))
    lap("setInputTarget (kNN cov, %d pts)" % len(ds), lambda: g.setInputTarget(ds))
    lap("setInputSource (kNN cov, %d pts)" % len(scans[0]), lambda: g.setInputSource(scans[0]))
    lap("align", lambda: g.align())
    print("  iterations %d" % g.result.iterations)
    lap("getFitnessScore", lambda: g.getFitnessScore())


if __name__ == "__main__":
    main()
