"""Experiment: do H2D copies overlap our kernels?  (1) torch-only copy vs concurrent kernels, (2) per-call wall
times of the prefetch / frames_cloud / downsample sequence."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from online_3d_reconstruction_b200 import abi
from online_3d_reconstruction_b200.pose import Pose

def torch_overlap():
    n = 184_320_000
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    a = torch.empty(256 << 20, dtype=torch.float32, device="cuda")
    b = torch.empty_like(a)
    sc, sk = torch.cuda.Stream(), torch.cuda.Stream()
    def copy_ms(concurrent):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if concurrent:
            with torch.cuda.stream(sk):
                for _ in range(20): b.copy_(a)
        with torch.cuda.stream(sc):
            e0.record(); d.copy_(h, non_blocking=True); e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    for c in (False, True, False, True):
        print(f"torch H2D 184 MB, concurrent kernels={c}: {copy_ms(c):.2f} ms", flush=True)

def api_timeline(prefetch, K=6):
    wl = "config2_semidense_720p"
    rows, cols, dt, J, v, mp, nd, F, seed, qs = bench.WORKLOADS[wl]
    disp, bgr, T = bench.make_data(wl, 0, 1, K + 1)
    P = Pose(bench.params_for(wl, 0))
    hd = [torch.from_numpy(a).pin_memory() for a in disp]; hb = [torch.from_numpy(a).pin_memory() for a in bgr]
    fr = [bench.frames_array([t.data_ptr() for t in hd], disp[0].strides[0], [t.data_ptr() for t in hb], bgr[0].strides[0], T[s]) for s in range(K + 1)]
    out = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    rowsout = []
    if prefetch: P.prefetchCycle(fr[0], dt)
    for s in range(K):
        t0 = time.perf_counter()
        if prefetch: P.prefetchCycle(fr[s + 1], dt)
        t1 = time.perf_counter()
        P.createCycleClouds(fr[s], dt)
        t2 = time.perf_counter()
        P.downsamplePtCloud(out.numpy().view(abi.POINT))
        t3 = time.perf_counter()
        rowsout.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
    for r in rowsout: print(f"  prefetch={prefetch} chunkdev={os.environ.get('O3R_CHUNK_FRAMES_DEV')}: prefetch {r[0]:.2f} frames_cloud {r[1]:.2f} downsample {r[2]:.2f} total {sum(r):.2f} ms", flush=True)
    P.close()

if __name__ == "__main__":
    os.environ["O3R_TRACE"] = "1"
    api_timeline(False, 5); api_timeline(True, 5)
