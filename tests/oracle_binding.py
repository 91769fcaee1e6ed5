"""ctypes binding of oracle/_build/libo3r_oracle.so — TEST INFRASTRUCTURE (the checker).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from online_3d_reconstruction_b200 import abi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "libo3r_oracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", os.path.join(_ROOT, "oracle")], check=True,
                   stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        P, F = C.POINTER(abi.Params), C.POINTER(abi.Frame)
        vp, sz, szp = C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)
        L.orc_blur_u8.argtypes = [vp, sz, C.c_int, C.c_int, C.c_int, C.c_int, vp, sz]
        L.orc_sor.argtypes = [vp, sz, C.c_int, C.c_double, C.c_int, C.c_int, vp, vp]
        L.orc_bilateral_u8.argtypes = [vp, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, vp, sz]
        L.orc_create_single_img_pt_cloud.argtypes = [P, F, C.c_int, vp, sz, szp, vp, sz, szp]
        L.orc_transform_pt_cloud.argtypes = [vp, sz, C.POINTER(C.c_float), vp]
        L.orc_transform_pt_cloud.restype = None
        L.orc_voxel_grid.argtypes = [vp, sz, C.c_float, C.c_float, C.c_float, C.c_uint, vp, sz, szp,
                                     vp, vp, C.POINTER(C.c_int)]
        L.orc_downsample_pt_cloud.argtypes = [P, vp, sz, C.c_int, vp, sz, szp]
        L.orc_create_and_transform_pt_cloud.argtypes = [P, F, C.c_int, vp, sz, szp]
        L.orc_run_cycle.argtypes = [P, F, C.c_int, C.c_int, C.c_int, vp, sz, szp, vp]
        L.orc_generate_tmat.argtypes = [C.c_double] * 7 + [C.POINTER(C.c_float)]
        L.orc_mat4_mul.argtypes = [C.POINTER(C.c_float)] * 3
        L.orc_mat4_mul.restype = None
        L.orc_plane_fit.argtypes = [P, vp, sz, vp, sz, vp, C.POINTER(C.c_int), vp]
        L.orc_get_variance.argtypes = [P, vp, sz, C.c_int]
        L.orc_get_variance.restype = C.c_double
        L.orc_cell_key.argtypes = [C.c_float] * 6
        L.orc_cell_key.restype = C.c_uint64
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed: {rc}")


def blur_u8(src, kernel, mode):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    dst = np.empty_like(src)
    _check(lib().orc_blur_u8(src.ctypes.data, src.strides[0], src.shape[0], src.shape[1], kernel, mode,
                             dst.ctypes.data, dst.strides[0]), "blur")
    return dst


def bilateral_u8(src, d, sigma_color, sigma_space):
    """cv::bilateralFilter on a (rows, cols) or (rows, cols, 3) u8 image."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    cn = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty_like(src)
    _check(lib().orc_bilateral_u8(src.ctypes.data, src.strides[0], src.shape[0], src.shape[1], cn, d, float(sigma_color),
                                  float(sigma_space), dst.ctypes.data, dst.strides[0]), "bilateral")
    return dst


def sor(pts, mean_k=50, stddev_mul=1.0, threads=4, brute=False):
    """pcl::StatisticalOutlierRemoval -> (keep mask u8, mean neighbour distances f32)."""
    pts = np.ascontiguousarray(pts, dtype=abi.POINT)
    keep = np.zeros(max(1, pts.size), dtype=np.uint8)
    dist = np.zeros(max(1, pts.size), dtype=np.float32)
    _check(lib().orc_sor(pts.ctypes.data, pts.size, mean_k, float(stddev_mul), threads, int(brute), keep.ctypes.data,
                         dist.ctypes.data), "sor")
    return keep[:pts.size], dist[:pts.size]


def max_points(p, n_kp=0):
    ny, nx = abi.scan_dims(p)
    return ny * nx + (n_kp if p.jump_pixels != 1 else 0)


def create_single_img_pt_cloud(p, frame, disp_type, want_mask=False):
    cap = max_points(p, frame.n_kp)
    out = np.empty(cap, dtype=abi.POINT)
    ny, nx = abi.scan_dims(p)
    mask = np.zeros(ny * nx, dtype=np.uint8) if want_mask else None
    n, ns = C.c_size_t(0), C.c_size_t(0)
    _check(lib().orc_create_single_img_pt_cloud(C.byref(p), C.byref(frame), disp_type, out.ctypes.data, cap,
                                                C.byref(n), None if mask is None else mask.ctypes.data,
                                                0 if mask is None else mask.size, C.byref(ns)), "create")
    out = out[:n.value].copy()
    return (out, mask) if want_mask else out


def transform_pt_cloud(pts, T):
    pts = np.ascontiguousarray(pts, dtype=abi.POINT)
    out = np.empty_like(pts)
    Tc = (C.c_float * 16)(*np.asarray(T, dtype=np.float32).reshape(16))
    lib().orc_transform_pt_cloud(pts.ctypes.data, pts.size, Tc, out.ctypes.data)
    return out


def voxel_grid(pts, leaf, min_points=0):
    """-> (points, keys u64, counts u32, passthrough bool)"""
    pts = np.ascontiguousarray(pts, dtype=abi.POINT)
    n = pts.size
    out = np.empty(n, dtype=abi.POINT)
    keys = np.empty(n, dtype=np.uint64)
    counts = np.empty(n, dtype=np.uint32)
    m, pt = C.c_size_t(0), C.c_int(0)
    lx, ly, lz = (np.float32(v) for v in leaf)
    _check(lib().orc_voxel_grid(pts.ctypes.data, n, lx, ly, lz, min_points, out.ctypes.data, n, C.byref(m),
                                keys.ctypes.data, counts.ctypes.data, C.byref(pt)), "voxel_grid")
    k = m.value
    return out[:k].copy(), keys[:k].copy(), counts[:k].copy(), bool(pt.value)


def downsample_pt_cloud(p, pts, combined):
    pts = np.ascontiguousarray(pts, dtype=abi.POINT)
    out = np.empty(max(pts.size, 1), dtype=abi.POINT)
    m = C.c_size_t(0)
    _check(lib().orc_downsample_pt_cloud(C.byref(p), pts.ctypes.data, pts.size, int(combined), out.ctypes.data,
                                         out.size, C.byref(m)), "downsample")
    return out[:m.value].copy()


def create_and_transform_pt_cloud(p, frame, disp_type):
    cap = max(max_points(p, frame.n_kp), 1)
    out = np.empty(cap, dtype=abi.POINT)
    m = C.c_size_t(0)
    _check(lib().orc_create_and_transform_pt_cloud(C.byref(p), C.byref(frame), disp_type, out.ctypes.data, cap,
                                                   C.byref(m)), "create_and_transform")
    return out[:m.value].copy()


def run_cycle(p, frames, disp_type, threads, cloud_big=None, cloud_n=0):
    """Appends the cycle to cloud_big (allocated on first use) -> (cloud_big, cloud_n, frame_counts)."""
    n = len(frames)
    arr = (abi.Frame * n)(*frames)
    per = max(max_points(p, max(f.n_kp for f in frames)), 1)
    need = cloud_n + per * n
    if cloud_big is None or cloud_big.size < need:
        nb = np.empty(need, dtype=abi.POINT)
        if cloud_big is not None:
            nb[:cloud_n] = cloud_big[:cloud_n]
        cloud_big = nb
    cn = C.c_size_t(cloud_n)
    counts = np.zeros(n, dtype=np.uint32)
    _check(lib().orc_run_cycle(C.byref(p), arr, n, disp_type, threads, cloud_big.ctypes.data, cloud_big.size,
                               C.byref(cn), counts.ctypes.data), "run_cycle")
    return cloud_big, cn.value, counts


def generate_tmat(tx, ty, tz, qx, qy, qz, qw):
    out = (C.c_float * 16)()
    rc = lib().orc_generate_tmat(tx, ty, tz, qx, qy, qz, qw, out)
    if rc:
        raise ValueError("Exception: Sum of squares of quaternion values should be 1!")
    return np.array(out, dtype=np.float32).reshape(4, 4)


def mat4_mul(a, b):
    A = (C.c_float * 16)(*np.asarray(a, np.float32).reshape(16))
    B = (C.c_float * 16)(*np.asarray(b, np.float32).reshape(16))
    O = (C.c_float * 16)()
    lib().orc_mat4_mul(A, B, O)
    return np.array(O, dtype=np.float32).reshape(4, 4)


def plane_fit(p, labels, disp):
    labels = np.ascontiguousarray(labels, np.uint8)
    disp = np.ascontiguousarray(disp, np.uint8)
    coef = np.zeros((1023, 3), dtype=np.float64)
    img = np.zeros((p.rows, p.cols), dtype=np.float64)
    npl = C.c_int(0)
    _check(lib().orc_plane_fit(C.byref(p), labels.ctypes.data, labels.strides[0], disp.ctypes.data,
                               disp.strides[0], coef.ctypes.data, C.byref(npl), img.ctypes.data), "plane_fit")
    return coef[:npl.value].copy(), img


def get_variance(p, img, plane_fitted):
    img = np.ascontiguousarray(img, np.float64 if plane_fitted else np.uint8)
    return lib().orc_get_variance(C.byref(p), img.ctypes.data, img.strides[0], int(plane_fitted))


def cell_key(x, y, z, leaf):
    return lib().orc_cell_key(np.float32(x), np.float32(y), np.float32(z), *(np.float32(v) for v in leaf))
