import os, sys
os.environ["O3R_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench, torch
from online_3d_reconstruction_b200.pose import Pose
wl = "config2_semidense_720p_sor"
rows, cols, dt, J, v, mp, nd, F, seed, qs = bench.WORKLOADS[wl]
disp, bgr, T = bench.make_data(wl, 0, 1, 1)
P = Pose(bench.params_for(wl, 0))
dev = torch.device("cuda", 0)
dd = [torch.from_numpy(a).to(dev) for a in disp]; db = [torch.from_numpy(a).to(dev) for a in bgr]
fr = bench.frames_array([t.data_ptr() for t in dd], disp[0].strides[0], [t.data_ptr() for t in db], bgr[0].strides[0], T[0])
P.createCycleClouds(fr, dt, device_pointers=True)
P.close()
