import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from online_3d_reconstruction_b200 import abi, synth
from online_3d_reconstruction_b200.pose import Pose
keep=[]
p = abi.make_params(jump_pixels=1, voxel_size=0.05, merge_mode=abi.MERGE_ACCUMULATE_FUSED)
seq = synth.sequence(1002, int(sys.argv[1]), 720, 1280)
fr=[abi.make_frame(d,i,T,keep=keep) for d,i,T in seq]
with Pose(p) as P:
    c=P.createCycleClouds(fr)
    print("engine", P.lastCycleEngine(), c)
