"""Host-side pose -> 4x4 float matrix composition (the input the hot path consumes).

generate_tmat follows generateTmat (pose_functions.cpp:1178-1356): quaternion -> 3x3 double, transposed,
stored into float 4x4 factors, product t_wh*r_wh*r_invert_y*r_flip_xy*t_hi*r_invert_i*r_yi*r_xi evaluated
left to right in float (:1341).  mat4_mul is the float 4x4 product used for T_SVD * t_mat (pose.cpp:232)
and tf_icp * t_mat (pose.cpp:318-319).  Runs once per frame on the host, as in the reference.
"""
import math

import numpy as np

# pose.h:142-147
TRANS_HI = (-0.300, -0.040, -0.350)
THETA_XI = -1.1408 * 3.141592653589793238463 / 180
THETA_YI = 1.1945 * 3.141592653589793238463 / 180


def mat4_mul(a, b):
    """float32 4x4 product, each entry accumulated left to right (k = 0..3) in float32."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    out = np.empty((4, 4), dtype=np.float32)
    for i in range(4):
        s = a[i, 0] * b[0, :]
        for k in range(1, 4):
            s = (s + a[i, k] * b[k, :]).astype(np.float32)
        out[i, :] = s
    return out


def generate_tmat(tx, ty, tz, qx, qy, qz, qw):
    f32 = np.float32
    r_xi = np.eye(4, dtype=f32)
    r_xi[1, 1] = math.cos(THETA_XI); r_xi[1, 2] = -math.sin(THETA_XI)
    r_xi[2, 1] = math.sin(THETA_XI); r_xi[2, 2] = math.cos(THETA_XI)
    r_yi = np.eye(4, dtype=f32)
    r_yi[0, 0] = math.cos(THETA_YI); r_yi[0, 2] = math.sin(THETA_YI)
    r_yi[2, 0] = -math.sin(THETA_YI); r_yi[2, 2] = math.cos(THETA_YI)
    r_invert_i = np.diag([1, -1, -1, 1]).astype(f32)
    r_invert_y = np.diag([1, -1, 1, 1]).astype(f32)
    t_hi = np.eye(4, dtype=f32)
    t_hi[0, 3], t_hi[1, 3], t_hi[2, 3] = TRANS_HI
    r_flip_xy = np.zeros((4, 4), dtype=f32)
    r_flip_xy[3, 3] = 1; r_flip_xy[1, 0] = 1; r_flip_xy[0, 1] = 1; r_flip_xy[2, 2] = 1
    sqw, sqx, sqy, sqz = qw * qw, qx * qx, qy * qy, qz * qz
    if sqw + sqx + sqy + sqz < 0.99 or sqw + sqx + sqy + sqz > 1.01:  # pose_functions.cpp:1279-1280
        raise ValueError("Exception: Sum of squares of quaternion values should be 1! i.e., quaternion should be "
                         "homogeneous!")
    rot = np.zeros((3, 3))
    rot[0, 0] = sqx - sqy - sqz + sqw
    rot[1, 1] = -sqx + sqy - sqz + sqw
    rot[2, 2] = -sqx - sqy + sqz + sqw
    t1, t2 = qx * qy, qz * qw
    rot[0, 1] = 2.0 * (t1 + t2); rot[1, 0] = 2.0 * (t1 - t2)
    t1, t2 = qx * qz, qy * qw
    rot[0, 2] = 2.0 * (t1 - t2); rot[2, 0] = 2.0 * (t1 + t2)
    t1, t2 = qy * qz, qx * qw
    rot[1, 2] = 2.0 * (t1 + t2); rot[2, 1] = 2.0 * (t1 - t2)
    r_wh = np.eye(4, dtype=f32)
    r_wh[:3, :3] = rot.T.astype(f32)
    t_wh = np.eye(4, dtype=f32)
    t_wh[0, 3], t_wh[1, 3], t_wh[2, 3] = tx, ty, tz
    m = mat4_mul(t_wh, r_wh)
    for f in (r_invert_y, r_flip_xy, t_hi, r_invert_i, r_yi, r_xi):
        m = mat4_mul(m, f)
    return m
