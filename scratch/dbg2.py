import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from online_3d_reconstruction_b200 import abi, synth
from test_gpu_fused import _run_cycles, FUSED
from test_gpu_parity import SMALL4, _frames, _close
keep=[]
geom=SMALL4
for v in (0.08, 0.1, 0.12):
    p = abi.make_params(jump_pixels=1, voxel_size=v, merge_mode=FUSED, **geom)
    cycles = [_frames(290, 3, geom["rows"], geom["cols"], keep=keep)]
    try:
        got, exp = _run_cycles(p, cycles, expect_engine=1)
        _close(got, exp)
        print(v, "ok", len(got))
    except AssertionError as e:
        print(v, "FAIL", str(e)[:300])
