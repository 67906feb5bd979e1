"""Runs the independent numpy / scipy restatement (independent_restatement.py) on the bundled Velodyne pair and freezes
its results in tests/golden/independent_restatement.json.  tests/test_oracle_cpu.py asserts that the C++ oracle reproduces
them (poses and scores at 1e-6, counts exactly): the second, independent pin of the oracle that the reference's own tests
(a 0.05 m / 1 deg band, gicp_test.cpp:147-201) do not give.

    python tests/golden/make_reference_fixtures.py      (needs only numpy / scipy and tests/golden/velodyne_pair.npz)
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import independent_restatement as R  # noqa: E402

NDT_CASES = {
    # name: (voxel-grid leaf of both clouds, resolution, eps, max_iter, guess translation / yaw)
    "product_identity": (0.2, 1.0, 0.01, 64, None),
    "product_offset_guess": (0.2, 1.0, 0.01, 64, (0.3, -0.2, 0.05, 0.03)),
    "res2_identity": (0.2, 2.0, 0.01, 64, None),
}


def guess_matrix(g):
    if g is None:
        return None
    T = np.eye(4, dtype=np.float32)
    c, s = np.cos(g[3]), np.sin(g[3])
    T[:2, :2] = [[c, -s], [s, c]]
    T[:3, 3] = g[:3]
    return T


def main():
    z = np.load(os.path.join(HERE, "velodyne_pair.npz"))
    out = {"ndt": {}, "fast_gicp": {}}
    clouds = {}
    for leaf in (0.2,):
        t0 = time.time()
        clouds[leaf] = (R.voxel_grid(z["target"], leaf), R.voxel_grid(z["source"], leaf))
        print("voxel grid %.1f: %d / %d points (%.1f s)" % (leaf, len(clouds[leaf][0]), len(clouds[leaf][1]), time.time() - t0))
    out["voxel_grid"] = {"0.2": dict(n_target=int(len(clouds[0.2][0])), n_source=int(len(clouds[0.2][1])),
                                     target_checksum=float(clouds[0.2][0].astype(np.float64).sum()),
                                     source_checksum=float(clouds[0.2][1].astype(np.float64).sum()))}
    for name, (leaf, res, eps, it, g) in NDT_CASES.items():
        td, sd = clouds[leaf]
        n = R.NDT(resolution=res, step_size=0.1, trans_eps=eps, max_iter=it)
        t0 = time.time()
        n.set_target(td)
        n.set_source(sd)
        G = guess_matrix(g)
        n.align(G)
        # derivatives at a fixed pose: an independent check of score / gradient / Hessian themselves
        p = np.array([0.25, -0.2, 0.03, 0.004, -0.006, 0.02])
        sc, gr, H, terms = n.derivatives(p, R.convert_transform(p))
        out["ndt"][name] = dict(leaf=leaf, resolution=res, eps=eps, max_iter=it, guess=None if g is None else list(g),
                                iterations=n.iterations, converged=bool(n.converged), evaluations=n.stats["evals"], trials=n.stats["trials"],
                                hessian_recomputes=n.stats["hess"], trans_probability=n.trans_probability,
                                T=[float(v) for v in n.final_T.ravel()], n_valid_voxels=int(n.grid.valid.sum()), n_voxels=int(len(n.grid.valid)),
                                fitness=R.fitness(sd, td, n.final_T),
                                probe=dict(p=p.tolist(), score=sc, gradient=gr.tolist(), hessian=H.ravel().tolist(), terms=terms))
        print("ndt %s: %d iterations, %d evaluations, %d trials, %d hessians (%.1f s)" %
              (name, n.iterations, n.stats["evals"], n.stats["trials"], n.stats["hess"], time.time() - t0))
    td, sd = clouds[0.2]
    for name, kw, g in (("gtest_recipe", {}, None), ("node_config", dict(max_corr=2.0, max_iter=100, trans_eps=0.01), (0.3, -0.2, 0.05, 0.03))):
        t0 = time.time()
        f = R.FastGICP(**kw)
        f.set_target(td)
        f.set_source(sd)
        f.align(None if g is None else guess_matrix(g).astype(np.float64))
        out["fast_gicp"][name] = dict(params={k: float(v) for k, v in kw.items()}, guess=None if g is None else list(g), iterations=f.iterations,
                                      converged=bool(f.converged), linearize_calls=f.stats["linearize"], compute_error_calls=f.stats["compute_error"],
                                      T=[float(v) for v in f.final_T.ravel()], fitness=R.fitness(sd, td, f.final_T),
                                      cov_checksum=float(np.abs(f.src_cov).sum()), cov_first=f.src_cov[0].ravel().tolist())
        print("fast_gicp %s: nr_iterations %d, %d linearisations (%.1f s)" % (name, f.iterations, f.stats["linearize"], time.time() - t0))
    with open(os.path.join(HERE, "independent_restatement.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
