import sys, numpy as np
sys.path.insert(0, "/root/repo")
from lidar_graph_slam_b200 import api, synth
n_pairs = 4096
d = synth.loop_keyframes(n_pairs=n_pairs, n_keyframes=41, n_azimuth=900, n_unique=2)
kf = api.KeyFrameArray()
for c, P in zip(d["clouds"], d["poses"]):
    kf.push(c, P)
import time
t0 = time.time()
recs = kf.batch_align(d["scan_ids"], d["center_ids"], search_key_frame_num=20, n_workers=4)
print("pairs/s", n_pairs / (time.time() - t0))
te, re, fit = [], [], []
for r, corr in zip(recs, d["corrections"]):
    T = np.array(r.T, np.float32).reshape(4, 4, order="F").astype(np.float64)
    E = np.linalg.inv(corr) @ T
    te.append(np.linalg.norm(E[:3, 3])); re.append(np.degrees(np.arccos(np.clip((np.trace(E[:3, :3]) - 1) / 2, -1, 1)))); fit.append(r.fitness)
te, re, fit = np.array(te), np.array(re), np.array(fit)
print("t err quantiles 50/90/99/max", np.quantile(te, [0.5, 0.9, 0.99, 1.0]))
print("r err deg quantiles", np.quantile(re, [0.5, 0.9, 0.99, 1.0]))
print("bad (>0.1 m):", (te > 0.1).sum(), "fitness of bad:", np.sort(fit[te > 0.1])[:10], "fitness of good max:", fit[te <= 0.1].max(), "median", np.median(fit))
print("iterations max", max(r.iterations for r in recs), "converged", sum(r.converged for r in recs))
