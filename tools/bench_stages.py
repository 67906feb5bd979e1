#!/usr/bin/env python
"""Per-stage device timings on the BASELINE configs (run on the GPU box), each against its HBM roofline:
algorithmic bytes / CUDA-event time / MEASURED_PEAKS.json:hbm_gbs.

    python tools/bench_stages.py [--reps 20] [--big] > gpurun_out/stages.json

Stages: cfg 1 prefilter (262 144-pt sweep, leaf 0.2 / 0.1, range crop 1.0), NDT target build at 1 M (cfg 0) and 20 M
points (cfg 3, 1.0 / 0.5 m, --big), NDT align against both, the device radix sort alone, GICP covariances / align /
fitness on cfg 2 sweeps.  One JSON object per line.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    import torch
    from lidar_graph_slam_b200 import api, synth
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = api.Context(0, stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, reps=args.reps, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
        b = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
        for i in range(reps):
            flush.zero_()
            a[i].record(stream)
            fn()
            b[i].record(stream)
        torch.cuda.synchronize()
        ms = sorted(x.elapsed_time(y) for x, y in zip(a, b))
        return ms[len(ms) // 2], ms[0]

    def emit(stage, ms_med, ms_min, alg_bytes, **kw):
        gbs = alg_bytes / (ms_med * 1e-3) / 1e9
        print(json.dumps(dict(stage=stage, ms_median=ms_med, ms_min=ms_min, algorithmic_bytes=alg_bytes, gbs=gbs, frac_of_hbm_peak=gbs / peak, **kw)), flush=True)

    # ---- cfg 1: prefilter
    sweeps = synth.prefilter_sweeps(n_sweeps=1)
    sw = torch.from_numpy(sweeps[0]).cuda()
    for leaf in (0.2, 0.1):
        vg = api.VoxelGrid(ctx)
        vg.setLeafSize(leaf)
        vg.setRangeCrop(1.0)
        vg.setInputCloud(sw)
        out = vg.filter()
        n, v = sw.shape[0], out.shape[0]
        l0 = ctx.launch_count
        med, mn = timed(lambda: vg.filter())
        emit("cfg1 prefilter leaf %.1f (device-resident sweep)" % leaf, med, mn, 16 * n + 4 * n + 4 * n + 16 * v, n=n, voxels=v,
             sweeps_per_s=1e3 / med, launches_per_call=(ctx.launch_count - l0) / (args.reps + 3))
        host = torch.from_numpy(sweeps[0]).pin_memory().numpy()
        vg.setInputCloud(host)
        med, mn = timed(lambda: vg.filter())
        emit("cfg1 prefilter leaf %.1f (host sweep in, host outputs back)" % leaf, med, mn, 16 * n + 4 * n + 4 * n + 16 * v, n=n, voxels=v, sweeps_per_s=1e3 / med)

    # ---- the radix sort alone
    for n, bits in ((262144, 21), (1_000_000, 21), (20_000_000, 24)):
        if n > 1_000_000 and not args.big:
            continue
        rs = np.random.RandomState(1)
        keys = rs.randint(0, 1 << bits, n).astype(np.uint32)
        # the C ABI sort hook works on host arrays; time the device part through the voxel grid instead (above) and
        # report the hook's end-to-end time here for reference
        med, mn = timed(lambda: api.sort_pairs(keys, np.arange(n, dtype=np.uint32), bits, ctx), reps=5, warm=1)
        emit("radix sort hook %d pairs, %d bits (host in/out, includes 4 PCIe copies)" % (n, bits), med, mn, ((bits + 7) // 8) * 16 * n + 4 * n, n=n)

    # ---- cfg 0 / cfg 3: NDT target build + align
    cases = [("cfg0 1M map", synth.ndt_scan_to_map(), (1.0,))]
    if args.big:
        cases.append(("cfg3 20M map", synth.rolling_map(), (1.0, 0.5)))
    for name, d, ress in cases:
        tgt = torch.from_numpy(d["target"]).cuda()
        src = torch.from_numpy(d["source"]).cuda()
        for res in ress:
            ndt = api.NormalDistributionsTransform(ctx)
            ndt.setResolution(res)
            ndt.setStepSize(0.1)
            ndt.setTransformationEpsilon(0.01)
            ndt.setMaximumIterations(64)
            ndt.setInputTarget(tgt)
            gi = ndt.grid_info()
            l0 = ctx.launch_count
            med, mn = timed(lambda: ndt.setInputTarget(tgt), reps=max(3, args.reps // 4))
            emit("%s: NDT target build res %.1f (device-resident map)" % (name, res), med, mn, 16 * tgt.shape[0] + 48 * int(gi.n_voxels), n=tgt.shape[0],
                 voxels=int(gi.n_voxels), valid=int(gi.n_valid), dense=bool(gi.dense), launches_per_call=(ctx.launch_count - l0) / (max(3, args.reps // 4) + 3))
            ndt.setInputSource(src)
            ndt.align(d["guess"])
            ev = ndt.result.evaluations + ndt.result.hessian_recomputes
            ndt.profile(True)
            ndt.align(d["guess"])
            prof = ndt.profile(False)
            per_eval = src.shape[0] * (16 + 7 * 8) + 40.0 * prof["terms_last_eval"]
            med, mn = timed(lambda: ndt.align(d["guess"]))
            emit("%s: NDT align res %.1f" % (name, res), med, mn, per_eval * ev, evaluations=ev, iterations=int(ndt.result.iterations),
                 aligns_per_s=1e3 / med, h_bar=prof["terms_last_eval"] / src.shape[0])
            med, mn = timed(lambda: ndt.getFitnessScore())
            emit("%s: getFitnessScore (1-NN of %d in %d)" % (name, src.shape[0], tgt.shape[0]), med, mn, 32 * src.shape[0])
            del ndt

    # ---- cfg 2: GICP scan-to-scan
    seq, _poses = synth.odometry_sequence(n_sweeps=3)
    vg = api.VoxelGrid(ctx)
    vg.setLeafSize(0.25)
    clouds = []
    for sweep in seq[:2]:
        vg.setInputCloud(synth.drop_invalid(sweep))
        clouds.append(torch.from_numpy(vg.filter(want_membership=False)).cuda())
    g = api.FastGICP(ctx)
    g.setMaxCorrespondenceDistance(1.0)
    g.setInputTarget(clouds[0])
    g.setInputSource(clouds[1])
    g.align()
    n_s, n_t = clouds[1].shape[0], clouds[0].shape[0]

    def gicp_full():
        g.setInputTarget(clouds[0])
        g.setInputSource(clouds[1])
        g.align()
    med, mn = timed(gicp_full)
    it = int(g.result.iterations) + 1
    emit("cfg2 GICP scan-to-scan: set target + set source + align (kNN-20 covariances of both clouds, %d LM iterations)" % it, med, mn,
         (16 + 24) * (n_s + n_t) + it * n_s * (16 + 24 + 16 + 24), n_source=n_s, n_target=n_t, pairs_per_s=1e3 / med)
    med, mn = timed(lambda: g.getFitnessScore())
    emit("cfg2 GICP getFitnessScore", med, mn, 32 * n_s)

    # ---- PointCloud2 ingest: Velodyne PointXYZIRT payload, point_step 22
    n = 262144
    rs = np.random.RandomState(3)
    raw = rs.randint(0, 256, size=(n, 22)).astype(np.uint8)
    raw[:, 0:16] = np.ascontiguousarray(sweeps[0][:n]).view(np.uint8).reshape(n, 16)
    msg = raw.tobytes()
    fields = {"x": (0, 7), "y": (4, 7), "z": (8, 7), "intensity": (12, 7), "ring": (16, 4), "time": (18, 7)}
    med, mn = timed(lambda: api.from_pointcloud2(msg, n, 1, 22, fields, ctx=ctx))
    emit("PointCloud2 ingest, point_step 22 (pageable host payload in, device xyzi out; H2D of %d bytes included)" % (n * 22), med, mn, n * (22 + 16), n=n,
         sweeps_per_s=1e3 / med)

    # ---- key-frame array: sub-map assembly: 20 key frames of ~50 k points, newest first, then + VoxelGrid 0.5
    kf = api.KeyFrameArray(ctx)
    for i in range(24):
        P = np.eye(4, dtype=np.float32)
        P[:3, 3] = [1.0 * i, 0.1 * i, 0]
        kf.push(sw[i * 1000: i * 1000 + 50000], P)
    ids = [23 - k for k in range(20)]
    import ctypes as C
    out_p, out_n = C.c_void_p(), C.c_int64()
    idarr = np.asarray(ids, np.int32)

    def assemble(leaf):
        api.check(kf._L.lgs_keyframes_assemble(kf._h, idarr.ctypes.data_as(C.c_void_p), len(ids), float(leaf), C.byref(out_p), C.byref(out_n)))
    med, mn = timed(lambda: assemble(0.0))
    emit("sub-map assembly: 20 key frames x 50 000 points (transform + concatenate, device resident)", med, mn, 32 * 20 * 50000, n=20 * 50000)
    # rolling map: a key-frame change as full rebuild (assemble + setInputTarget) and as incremental update
    nd_full, nd_inc = api.NormalDistributionsTransform(ctx), api.NormalDistributionsTransform(ctx)
    for x in (nd_full, nd_inc):
        x.setResolution(1.0)
    state = {"k": 0}

    def window():
        state["k"] = (state["k"] + 1) % 4
        return np.asarray([23 - state["k"] - j for j in range(20)], np.int32)

    def full_rebuild():
        w = window()
        api.check(kf._L.lgs_keyframes_assemble(kf._h, w.ctypes.data_as(C.c_void_p), len(w), 0.0, C.byref(out_p), C.byref(out_n)))
        api.check(nd_full._L.lgs_ndt_set_target_dev(nd_full._h, out_p, out_n.value))
    med, mn = timed(full_rebuild)
    emit("rolling map, key-frame change: assemble 20 x 50 000 points + NDT target build (full rebuild)", med, mn, 32 * 20 * 50000 + 16 * 20 * 50000, n=20 * 50000)
    med, mn = timed(lambda: nd_inc.setInputTargetKeyFrames(kf, window()))
    emit("rolling map, key-frame change: lgs_ndt_set_target_keyframes (1 of 20 key frames voxelised, cached partial sums merged)", med, mn,
         32 * 50000 + 16 * 50000 + 88 * int(nd_inc.grid_info().n_voxels) * 4, n=20 * 50000)
    med, mn = timed(lambda: assemble(0.5))
    emit("sub-map assembly + VoxelGrid 0.5 m (GBS:297-313)", med, mn, 32 * 20 * 50000 + (16 + 8) * 20 * 50000 + 16 * int(out_n.value), n=20 * 50000, voxels=int(out_n.value))

    # ---- the node's other two registration methods on the cfg 2 clouds
    icp = api.IterativeClosestPoint(ctx)
    icp.setMaxCorrespondenceDistance(30)
    icp.setMaximumIterations(100)
    icp.setTransformationEpsilon(1e-8)
    icp.setEuclideanFitnessEpsilon(1e-6)
    icp.setInputTarget(clouds[0])
    icp.setInputSource(clouds[1])
    icp.align()
    it = int(icp.result.iterations)
    med, mn = timed(lambda: icp.align())
    emit("cfg2 clouds, pcl::IterativeClosestPoint (GBS:142-151 settings): align, %d iterations" % it, med, mn, it * n_s * 48, n_source=n_s, n_target=n_t,
         aligns_per_s=1e3 / med, us_per_iteration=1e3 * med / max(it, 1))
    go = api.GeneralizedIterativeClosestPoint(ctx)
    go.setMaxCorrespondenceDistance(2.0)
    go.setMaximumIterations(100)
    go.setTransformationEpsilon(0.01)
    go.setInputTarget(clouds[0])
    go.setInputSource(clouds[1])
    go.align()
    nf = int(go.result.evaluations + go.result.line_search_trials)
    med, mn = timed(lambda: go.align())
    emit("cfg2 clouds, pclomp GICP (BFGS): align with kept covariances, %d outer iterations, %d functor evaluations" % (int(go.result.iterations), nf), med, mn,
         nf * n_s * (16 + 16 + 36 + 4), n_source=n_s, n_target=n_t, aligns_per_s=1e3 / med, us_per_functor_evaluation=1e3 * med / max(nf, 1))


if __name__ == "__main__":
    main()
