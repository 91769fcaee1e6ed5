/*
 * o3r_oracle.h — CPU restatement of the reference's reconstruction hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under online_3d_reconstruction_b200/ (the product) may
 * include, link, import or execute this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do, and only as the checker / reported baseline.
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).
 * The arithmetic of pcl::VoxelGrid / pcl::transformPointCloud / cv::Mat_ ops lives in PCL 1.8 and
 * OpenCV 3.1.0, which are NOT vendored in the reference and not installed here; those parts are
 * restated from the libraries' published algorithms (see DESIGN.md "Oracle" for what pins them).
 *
 * Build: make -C oracle   (g++ -O2 -ffp-contract=off; IEEE f32/f64, no FMA contraction).
 */
#ifndef O3R_ORACLE_H
#define O3R_ORACLE_H

#include "../include/o3r.h"

#ifdef __cplusplus
extern "C" {
#endif

/* cv::medianBlur / cv::blur on one u8 plane (the blur step, pose_functions.cpp:1040-1047 with the
 * north-star filter choice).  median: odd k, BORDER_REPLICATE; box: anchor k/2, BORDER_REFLECT_101,
 * round-half-even(S/k^2). */
int orc_blur_u8(const uint8_t* src, size_t src_step, int rows, int cols, int kernel, int mode,
                uint8_t* dst, size_t dst_step);

/* cv::bilateralFilter on a u8 image with cn = 1 or 3 interleaved channels (OpenCV 3.1 bilateralFilter_8u);
 * used to pin the model on the reference's own build/output/bilateralFiltered_15.png (cn = 3). */
int orc_bilateral_u8(const uint8_t* src, size_t src_step, int rows, int cols, int cn, int d, double sigma_color,
                     double sigma_space, uint8_t* dst, size_t dst_step);

/* createSingleImgPtCloud  pose_functions.cpp:1030-1134.  Camera-frame points (before transform).
 * mask (optional) gets 1 byte per scanned grid sample, *n_scanned the number of samples. */
int orc_create_single_img_pt_cloud(const o3r_params* p, const o3r_frame* f, int disp_type,
                                   o3r_point* out, size_t cap, size_t* n_out,
                                   uint8_t* mask, size_t mask_cap, size_t* n_scanned);

/* transformPtCloud  pose_functions.cpp:1358-1362 -> pcl::transformPointCloud (dense branch);
 * same formula in-tree at transformPoint pose_functions.cpp:1830-1832.  in == out allowed. */
void orc_transform_pt_cloud(const o3r_point* in, size_t n, const float T[16], o3r_point* out);

/* pcl::VoxelGrid<PointXYZRGB>::applyFilter (PCL 1.8, call sites pose_functions.cpp:1689-1700). */
int orc_voxel_grid(const o3r_point* pts, size_t n, float lx, float ly, float lz, unsigned min_points,
                   o3r_point* out, size_t cap, size_t* n_out,
                   uint64_t* keys, uint32_t* counts, int* passthrough);

/* downsamplePtCloud  pose_functions.cpp:1654-1709 (StatisticalOutlierRemoval :1673-1686 EXCLUDED,
 * SURVEY §8f-1). */
/* pcl::StatisticalOutlierRemoval (pose_functions.cpp:1673-1686): keep[i] = 0 for removed points; dist_out (optional)
 * receives every point's mean distance to its mean_k nearest neighbours.  brute = 1: all-pairs search (small clouds). */
int orc_sor(const o3r_point* pts, size_t n, int mean_k, double stddev_mul, int threads, int brute,
            uint8_t* keep, float* dist_out);

int orc_downsample_pt_cloud(const o3r_params* p, const o3r_point* pts, size_t n, int combined,
                            o3r_point* out, size_t cap, size_t* n_out);

/* createAndTransformPtCloud  pose.cpp:596-636. */
int orc_create_and_transform_pt_cloud(const o3r_params* p, const o3r_frame* f, int disp_type,
                                      o3r_point* out, size_t cap, size_t* n_out);

/* One cycle on `threads` workers (pose.cpp:361-434: fan-out, join, ordered concat, append), then
 * optionally the final combined downsample (pose.cpp:527-536).  cloud_big: caller buffer that the
 * cycle's clouds are appended to at *cloud_n (capacity cloud_cap).  Returns 0 / <0. */
int orc_run_cycle(const o3r_params* p, const o3r_frame* frames, int n, int disp_type, int threads,
                  o3r_point* cloud_big, size_t cloud_cap, size_t* cloud_n, uint32_t* frame_counts);

/* generateTmat  pose_functions.cpp:1178-1356: quaternion + translation -> float 4x4 (row-major out),
 * chain t_wh*r_wh*r_invert_y*r_flip_xy*t_hi*r_invert_i*r_yi*r_xi in float, left to right (:1341).
 * Returns -1 on the quaternion-norm throw (:1279-1280). */
int orc_generate_tmat(double tx, double ty, double tz, double qx, double qy, double qz, double qw,
                      float out[16]);

/* float 4x4 product out = a*b (pose.cpp:232, :318-319), row-major. */
void orc_mat4_mul(const float a[16], const float b[16], float out[16]);

/* createPlaneFittedDisparityImages  pose_functions.cpp:900-985.  coef_out: [1023][3] (a,b,c),
 * *n_planes = number of consecutive labels found; f64_out (optional) rows*cols doubles. */
int orc_plane_fit(const o3r_params* p, const uint8_t* labels, size_t labels_step,
                  const uint8_t* disp, size_t disp_step, double* coef_out, int* n_planes,
                  double* f64_out);

/* getMean / getVariance  pose_functions.cpp:987-1028 (plane_fitted: image is f64). */
double orc_get_variance(const o3r_params* p, const void* img, size_t step, int plane_fitted);

/* 64-bit absolute cell key of one point on a grid (SURVEY §8a row VG). */
uint64_t orc_cell_key(float x, float y, float z, float lx, float ly, float lz);

size_t orc_sizeof(int which);

#ifdef __cplusplus
}
#endif
#endif
