"""Deterministic synthetic UAV sequences at the reference frame geometry.

uav720(seed, n) / uav4k(seed, n) return disparity, BGR image and per-frame float 4x4 matrices whose value
distributions follow the reference's fixtures (disparity mean ~108, ROI variance ~2.3, range 65..127,
1.5 % dropouts, blank left eighth; 0.45 m between frames at 22 m altitude).
"""
import math

import numpy as np

from . import abi, tmat


def _disparity_field(rng, rows, cols):
    yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float32)
    s = np.zeros((rows, cols), np.float32)
    for _ in range(4):
        fx, fy = rng.uniform(0.5, 2.5, 2) * 2 * math.pi / np.array([cols, rows])
        ph = rng.uniform(0, 2 * math.pi)
        s += np.sin(fx * xx + fy * yy + ph).astype(np.float32)
    f = 108.0 + 6.0 * (s / 4.0) + rng.standard_normal((rows, cols), dtype=np.float32)
    for _ in range(3):  # obstacles
        h, w = rng.integers(rows // 12, rows // 4), rng.integers(cols // 16, cols // 5)
        y0, x0 = rng.integers(0, rows - h), rng.integers(0, cols - w)
        f[y0:y0 + h, x0:x0 + w] += rng.uniform(8, 20)
    return f


def make_frame_images(rng, rows, cols, disp_type=abi.DISP_U8):
    """-> (disparity array of disp_type, BGR u8 image)"""
    f = _disparity_field(rng, rows, cols)
    drop = rng.random((rows, cols), dtype=np.float32) < 0.015
    if disp_type == abi.DISP_U8:
        d = np.clip(np.rint(f), 0, 127).astype(np.uint8)
    elif disp_type == abi.DISP_U16:
        d = np.clip(np.rint(f * 200.0), 0, 127 * 200).astype(np.uint16)
    elif disp_type == abi.DISP_F32:
        d = np.clip(f, 0, 127).astype(np.float32)
    else:
        d = np.clip(f, 0, 127).astype(np.float64)
    d[drop] = 0
    d[:, :cols // 8] = 0
    blocks = rng.integers(0, 256, ((rows + 7) // 8, (cols + 7) // 8, 3), dtype=np.uint8)
    img = np.repeat(np.repeat(blocks, 8, axis=0), 8, axis=1)[:rows, :cols].astype(np.int16)
    img += np.rint(rng.standard_normal((rows, cols, 3), dtype=np.float32) * 4).astype(np.int16)
    return d, np.clip(img, 0, 255).astype(np.uint8)


def _quat(yaw, pitch, roll):
    cy, sy = math.cos(yaw / 2), math.sin(yaw / 2)
    cp, sp = math.cos(pitch / 2), math.sin(pitch / 2)
    cr, sr = math.cos(roll / 2), math.sin(roll / 2)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
            cr * cp * cy + sr * sp * sy)


def trajectory(rng, n, start=0, step=0.45, row_len=40, row_spacing=3.0, altitude=22.0):
    """Lawnmower sweep: per-frame float 4x4 matrices = T_correction * generateTmat(pose) composed in float32."""
    Ts = []
    for i in range(start, start + n):
        row, k = divmod(i, row_len)
        fwd = row % 2 == 0
        x = (k if fwd else row_len - 1 - k) * step
        y = row * row_spacing
        yaw = (0.0 if fwd else math.pi) + math.radians(rng.normal(0, 0.5))
        q = _quat(yaw, math.radians(rng.normal(0, 1.0)), math.radians(rng.normal(0, 1.0)))
        t_mav = tmat.generate_tmat(x, y, altitude, *q)
        corr = np.eye(4, dtype=np.float32)  # emulates T_SVD (pose.cpp:232): a small translation
        corr[:3, 3] = rng.normal(0, 0.15, 3)
        Ts.append(tmat.mat4_mul(corr, t_mav))
    return Ts


def sequence(seed, n, rows, cols, disp_type=abi.DISP_U8, n_images=None, traj_start=0):
    """n frames; only n_images distinct image pairs are generated (reused cyclically) when given."""
    rng = np.random.default_rng(seed)
    m = n if n_images is None else min(n, n_images)
    imgs = [make_frame_images(rng, rows, cols, disp_type) for _ in range(m)]
    Ts = trajectory(np.random.default_rng(seed + 7919), n, start=traj_start)
    return [(imgs[i % m][0], imgs[i % m][1], Ts[i]) for i in range(n)]


def uav720(seed, n, **kw):
    return sequence(seed, n, 720, 1280, **kw)


def uav4k(seed, n, disp_type=abi.DISP_U16, **kw):
    return sequence(seed, n, 2160, 3840, disp_type=disp_type, **kw)


def q_scaled(scale):
    """Q of cam13calib scaled for a `scale`x larger image (cx, cy, f scale; baseline term unchanged)."""
    q = list(abi.Q_CAM13)
    q[3] *= scale; q[7] *= scale; q[11] *= scale
    return tuple(q)
