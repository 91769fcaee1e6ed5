// tmat.hpp — host-side pose -> float 4x4 composition, the input the hot path consumes.
// generate_tmat follows generateTmat (pose_functions.cpp:1178-1356): quaternion -> 3x3 double (transposed), float
// 4x4 factors, product t_wh*r_wh*r_invert_y*r_flip_xy*t_hi*r_invert_i*r_yi*r_xi evaluated left to right in float.
#pragma once
#include <cmath>
#include <cstring>

namespace host {

struct Mat4 { float m[16]; };   // row-major

inline Mat4 mat4_mul(const Mat4& a, const Mat4& b) {   // pose.cpp:232, :318-319
    Mat4 o;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = a.m[i * 4] * b.m[j];
            s = s + a.m[i * 4 + 1] * b.m[4 + j];
            s = s + a.m[i * 4 + 2] * b.m[8 + j];
            s = s + a.m[i * 4 + 3] * b.m[12 + j];
            o.m[i * 4 + j] = s;
        }
    return o;
}

inline Mat4 mat4_identity() { Mat4 o; for (int i = 0; i < 16; ++i) o.m[i] = (i % 5 == 0) ? 1.f : 0.f; return o; }

// throws const char* like the reference (pose_functions.cpp:1279-1280)
inline Mat4 generate_tmat(double tx, double ty, double tz, double qx, double qy, double qz, double qw) {
    const double trans_x_hi = -0.300, trans_y_hi = -0.040, trans_z_hi = -0.350;   // pose.h:142-144
    const double PI = 3.141592653589793238463;
    const double theta_xi = -1.1408 * PI / 180, theta_yi = 1.1945 * PI / 180;     // pose.h:146-147
    auto zero = [] { Mat4 o; memset(o.m, 0, sizeof(o.m)); return o; };
    Mat4 r_xi = mat4_identity();
    r_xi.m[5] = (float)cos(theta_xi); r_xi.m[6] = (float)-sin(theta_xi);
    r_xi.m[9] = (float)sin(theta_xi); r_xi.m[10] = (float)cos(theta_xi);
    Mat4 r_yi = mat4_identity();
    r_yi.m[0] = (float)cos(theta_yi); r_yi.m[2] = (float)sin(theta_yi);
    r_yi.m[8] = (float)-sin(theta_yi); r_yi.m[10] = (float)cos(theta_yi);
    Mat4 r_invert_i = zero(); r_invert_i.m[15] = 1; r_invert_i.m[0] = 1; r_invert_i.m[5] = -1; r_invert_i.m[10] = -1;
    Mat4 r_invert_y = zero(); r_invert_y.m[15] = 1; r_invert_y.m[0] = 1; r_invert_y.m[5] = -1; r_invert_y.m[10] = 1;
    Mat4 t_hi = mat4_identity(); t_hi.m[3] = (float)trans_x_hi; t_hi.m[7] = (float)trans_y_hi; t_hi.m[11] = (float)trans_z_hi;
    Mat4 r_flip_xy = zero(); r_flip_xy.m[15] = 1; r_flip_xy.m[4] = 1; r_flip_xy.m[1] = 1; r_flip_xy.m[10] = 1;
    const double sqw = qw * qw, sqx = qx * qx, sqy = qy * qy, sqz = qz * qz;
    if (sqw + sqx + sqy + sqz < 0.99 || sqw + sqx + sqy + sqz > 1.01)
        throw "Exception: Sum of squares of quaternion values should be 1! i.e., quaternion should be homogeneous!";
    double rot[3][3];
    rot[0][0] = sqx - sqy - sqz + sqw; rot[1][1] = -sqx + sqy - sqz + sqw; rot[2][2] = -sqx - sqy + sqz + sqw;
    double t1 = qx * qy, t2 = qz * qw;
    rot[0][1] = 2.0 * (t1 + t2); rot[1][0] = 2.0 * (t1 - t2);
    t1 = qx * qz; t2 = qy * qw;
    rot[0][2] = 2.0 * (t1 - t2); rot[2][0] = 2.0 * (t1 + t2);
    t1 = qy * qz; t2 = qx * qw;
    rot[1][2] = 2.0 * (t1 + t2); rot[2][1] = 2.0 * (t1 - t2);
    Mat4 r_wh = zero(); r_wh.m[15] = 1;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r_wh.m[i * 4 + j] = (float)rot[j][i];   // rot = rot.t()
    Mat4 t_wh = mat4_identity(); t_wh.m[3] = (float)tx; t_wh.m[7] = (float)ty; t_wh.m[11] = (float)tz;
    Mat4 m = mat4_mul(t_wh, r_wh);   // pose_functions.cpp:1341, left to right
    m = mat4_mul(m, r_invert_y); m = mat4_mul(m, r_flip_xy); m = mat4_mul(m, t_hi);
    m = mat4_mul(m, r_invert_i); m = mat4_mul(m, r_yi); m = mat4_mul(m, r_xi);
    return m;
}

}  // namespace host
