import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from lidar_graph_slam_b200 import api
rs=np.random.RandomState(0)
for n_pts in (1000, 100000):
    pts=np.zeros((n_pts,4),np.float32); pts[:,:3]=rs.uniform(0.1,0.9,(n_pts,3))
    n=api.NormalDistributionsTransform(); n.setResolution(1.0)
    t=torch.from_numpy(pts).cuda()
    for _ in range(2): n.setInputTarget(t)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); n.setInputTarget(t); e1.record(); torch.cuda.synchronize()
    print(n_pts, "build ms", e0.elapsed_time(e1), "voxels", n.grid_info().n_voxels)
