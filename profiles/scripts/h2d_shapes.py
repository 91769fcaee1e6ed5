"""H2D microbenchmark behind the staging design: the same 150 MB of scan-ROI bytes of a 50-frame 720p cycle moved as
(a) 100 cudaMemcpy2DAsync calls (one per plane, two streams: what the library issues for separately allocated frames),
(b) 2 cudaMemcpy2DAsync calls over planes that sit back to back in one pinned arena (rows of all frames, margin rows included),
(c) 2 cudaMemcpy3DAsync calls whose extent is (ROI width, ROI rows, frames): the exact ROI bytes, (d) 100 contiguous copies of
whole row ranges.  Prints GB/s of ROI bytes."""
import time

import torch
from cuda.bindings import runtime as rt

F, ROWS, COLS, BB, X0 = 50, 720, 1280, 20, 160
W = COLS - BB - X0          # 1100
H = ROWS - 2 * BB           # 680
roi_bytes = F * H * W * 4


def chk(r):
    if isinstance(r, tuple):
        assert int(r[0]) == 0, r
        return r[1] if len(r) > 1 else None
    assert int(r) == 0, r


host_d = torch.empty(F * ROWS * COLS, dtype=torch.uint8).pin_memory()
host_c = torch.empty(F * ROWS * COLS * 3, dtype=torch.uint8).pin_memory()
dev_d = torch.empty(F * ROWS * COLS, dtype=torch.uint8, device="cuda")
dev_c = torch.empty(F * ROWS * COLS * 3, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
K = rt.cudaMemcpyKind.cudaMemcpyHostToDevice


def per_plane():
    for i in range(F):
        st = (s1 if i % 2 == 0 else s2).cuda_stream
        o = i * ROWS * COLS + BB * COLS + X0
        chk(rt.cudaMemcpy2DAsync(dev_d.data_ptr() + o, COLS, host_d.data_ptr() + o, COLS, W, H, K, st))
        o3 = i * ROWS * COLS * 3 + BB * COLS * 3 + X0 * 3
        chk(rt.cudaMemcpy2DAsync(dev_c.data_ptr() + o3, COLS * 3, host_c.data_ptr() + o3, COLS * 3, W * 3, H, K, st))


def arena():
    o = BB * COLS + X0
    rows = (F - 1) * ROWS + H
    chk(rt.cudaMemcpy2DAsync(dev_d.data_ptr() + o, COLS, host_d.data_ptr() + o, COLS, W, rows, K, s1.cuda_stream))
    chk(rt.cudaMemcpy2DAsync(dev_c.data_ptr() + 3 * o, COLS * 3, host_c.data_ptr() + 3 * o, COLS * 3, W * 3, rows, K, s2.cuda_stream))


def arena_groups(g):
    for a in range(0, F, g):
        n = min(g, F - a)
        o = a * ROWS * COLS + BB * COLS + X0
        rows = (n - 1) * ROWS + H
        st = (s1 if (a // g) % 2 == 0 else s2).cuda_stream
        chk(rt.cudaMemcpy2DAsync(dev_d.data_ptr() + o, COLS, host_d.data_ptr() + o, COLS, W, rows, K, st))
        chk(rt.cudaMemcpy2DAsync(dev_c.data_ptr() + 3 * o, COLS * 3, host_c.data_ptr() + 3 * o, COLS * 3, W * 3, rows, K, st))


def three_d():
    """ONE cudaMemcpy3DAsync per plane type: extent = (ROI width, ROI rows, frames) — the exact ROI bytes, no margin rows."""
    for (hp, dp, es, st) in ((host_d, dev_d, 1, s1), (host_c, dev_c, 3, s2)):
        prm = rt.cudaMemcpy3DParms()
        prm.srcPtr = rt.make_cudaPitchedPtr(hp.data_ptr(), COLS * es, COLS * es, ROWS)
        prm.dstPtr = rt.make_cudaPitchedPtr(dp.data_ptr(), COLS * es, COLS * es, ROWS)
        prm.srcPos = rt.make_cudaPos(X0 * es, BB, 0)
        prm.dstPos = rt.make_cudaPos(X0 * es, BB, 0)
        prm.extent = rt.make_cudaExtent(W * es, H, F)
        prm.kind = K
        chk(rt.cudaMemcpy3DAsync(prm, st.cuda_stream))


def rows_contig():
    for i in range(F):
        st = (s1 if i % 2 == 0 else s2).cuda_stream
        o = i * ROWS * COLS + BB * COLS
        chk(rt.cudaMemcpyAsync(dev_d.data_ptr() + o, host_d.data_ptr() + o, H * COLS, K, st))
        chk(rt.cudaMemcpyAsync(dev_c.data_ptr() + 3 * o, host_c.data_ptr() + 3 * o, H * COLS * 3, K, st))


def bench(name, fn, nbytes):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / 5)
    print(f"{name:40s} {best * 1e3:7.3f} ms  {roi_bytes / best / 1e9:6.1f} GB/s of ROI bytes ({nbytes / best / 1e9:6.1f} GB/s moved)")


rows_all = (F - 1) * ROWS + H
bench("100 x 2-D ROI copies, 2 streams", per_plane, roi_bytes)
bench("2 x 2-D copies over one arena", arena, rows_all * W * 4)
for g in (5, 10, 25):
    bench(f"2-D copies over groups of {g} frames", lambda g=g: arena_groups(g), 0)
bench("2 x 3-D copies (exact ROI of every frame)", three_d, roi_bytes)
bench("100 x contiguous row ranges", rows_contig, F * H * COLS * 4)
