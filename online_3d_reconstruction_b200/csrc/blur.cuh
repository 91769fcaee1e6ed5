// blur.cuh — the blur step of createSingleImgPtCloud (pose_functions.cpp:1040-1047) for u8 disparity,
// with the two filters north_star names: exact k x k median (cv::medianBlur: BORDER_REPLICATE, odd k)
// and normalised box (cv::blur: anchor k/2, BORDER_REFLECT_101, round-half-even(S / k^2)).
//
// One CTA produces a 128-row x 64-column block of outputs.  The input block with its halo is staged in
// shared memory once (the border rule is applied while staging), then each thread walks one row:
// it keeps a private 256-bin histogram (median; plus 16 coarse bins so the rank search is 32 probes)
// or a running sum (box) and slides it one column per step (k entries in, k out).  Outputs leave through
// a shared-memory tile so the global stores are row-coalesced.
#pragma once
#include "common.cuh"

namespace o3r {

constexpr int kBlurRows = 128;   // threads per CTA == output rows per CTA
constexpr int kBlurStrip = 64;   // output columns per CTA
constexpr int kBlurMaxK = 127;

struct BlurJob {
    const uint8_t* src; unsigned long long sstep;
    uint8_t* dst; unsigned long long dstep;
};

__host__ __device__ inline int blur_pitch(int k) {
    int w = kBlurStrip + k - 1;
    int p = (w + 3) / 4;
    if ((p & 1) == 0) ++p;  // odd number of words per row => the 32 rows of a warp hit 32 different banks
    return p * 4;
}
inline size_t blur_smem(int k, int mode) {
    size_t s = (size_t)(kBlurRows + k - 1) * blur_pitch(k) + (size_t)kBlurRows * kBlurStrip;
    s = (s + 15) & ~(size_t)15;
    if (mode == O3R_BLUR_MEDIAN) s += (size_t)(256 + 16) * kBlurRows * 2;
    return s;
}

__device__ __forceinline__ int dev_reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}

template <int MODE>
__global__ void __launch_bounds__(kBlurRows) k_blur(const BlurJob* __restrict__ jobs, int rows, int cols, int k,
                                                    int rx0, int ry0, int rx1, int ry1) {
    extern __shared__ __align__(16) unsigned char bsm[];
    const BlurJob job = jobs[blockIdx.z];
    const int a = k / 2, pitch = blur_pitch(k);
    const int bx = rx0 + blockIdx.x * kBlurStrip, by = ry0 + blockIdx.y * kBlurRows;
    const int tw = kBlurStrip + k - 1, th = kBlurRows + k - 1;
    unsigned char* tin = bsm;
    unsigned char* tout = bsm + (size_t)th * pitch;
    const int tid = threadIdx.x;
    // stage input block + halo, border rule applied here
    for (int i = tid; i < tw * th; i += kBlurRows) {
        const int ly = i / tw, lx = i - ly * tw;
        int gy = by - a + ly, gx = bx - a + lx;
        if (MODE == O3R_BLUR_MEDIAN) {
            gy = min(max(gy, 0), rows - 1);
            gx = min(max(gx, 0), cols - 1);
        } else {
            gy = dev_reflect101(gy, rows);
            gx = dev_reflect101(gx, cols);
        }
        tin[ly * pitch + lx] = job.src[(size_t)gy * job.sstep + gx];
    }
    __syncthreads();
    const unsigned char* myrow = tin + tid * pitch;   // window rows are myrow + dy*pitch, dy in [0,k)
    if (MODE == O3R_BLUR_MEDIAN) {
        uint16_t* hist = reinterpret_cast<uint16_t*>(bsm + (((size_t)th * pitch + (size_t)kBlurRows * kBlurStrip + 15) & ~(size_t)15));
        uint16_t* coarse = hist + 256 * kBlurRows;
        for (int b = 0; b < 256; ++b) hist[b * kBlurRows + tid] = 0;
        for (int b = 0; b < 16; ++b) coarse[b * kBlurRows + tid] = 0;
        for (int dy = 0; dy < k; ++dy)
            for (int dx = 0; dx < k; ++dx) {
                const int v = myrow[dy * pitch + dx];
                hist[v * kBlurRows + tid]++;
                coarse[(v >> 4) * kBlurRows + tid]++;
            }
        const int half = (k * k) / 2;
        for (int x = 0; x < kBlurStrip; ++x) {
            if (x > 0) {
                for (int dy = 0; dy < k; ++dy) {
                    const int vo = myrow[dy * pitch + x - 1], vn = myrow[dy * pitch + x + k - 1];
                    hist[vo * kBlurRows + tid]--; coarse[(vo >> 4) * kBlurRows + tid]--;
                    hist[vn * kBlurRows + tid]++; coarse[(vn >> 4) * kBlurRows + tid]++;
                }
            }
            int acc = 0, c = 0;
            for (; c < 16; ++c) {
                const int h = coarse[c * kBlurRows + tid];
                if (acc + h > half) break;
                acc += h;
            }
            int v = c * 16;
            for (;; ++v) {
                acc += hist[v * kBlurRows + tid];
                if (acc > half) break;
            }
            tout[tid * kBlurStrip + x] = (unsigned char)v;
        }
    } else {
        const int kk = k * k;
        int S = 0;
        for (int dy = 0; dy < k; ++dy)
            for (int dx = 0; dx < k; ++dx) S += myrow[dy * pitch + dx];
        for (int x = 0; x < kBlurStrip; ++x) {
            if (x > 0)
                for (int dy = 0; dy < k; ++dy) S += (int)myrow[dy * pitch + x + k - 1] - (int)myrow[dy * pitch + x - 1];
            int q = S / kk;
            const int rem = S - q * kk;
            if (2 * rem > kk) q += 1;
            else if (2 * rem == kk) q += (q & 1);
            tout[tid * kBlurStrip + x] = (unsigned char)min(q, 255);
        }
    }
    __syncthreads();
    for (int i = tid; i < kBlurRows * kBlurStrip; i += kBlurRows) {
        const int ly = i / kBlurStrip, lx = i - ly * kBlurStrip;
        const int gy = by + ly, gx = bx + lx;
        if (gy < ry1 && gx < rx1) job.dst[(size_t)gy * job.dstep + gx] = tout[i];
    }
}

}  // namespace o3r
