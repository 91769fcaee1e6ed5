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

// ---- cv::bilateralFilter(u8, d = k, sigmaColor = 2k, sigmaSpace = k/2): the reference's LIVE blur (pose_functions.cpp:1044)
// OpenCV 3.1 bilateralFilter_8u: BORDER_REFLECT_101 halo, circular tap mask of radius k/2 walked row-major, weight =
// space_weight[tap] * color_weight[|v - v0|] (float LUTs computed on the host with std::exp, exactly like OpenCV
// and the oracle do), two sequential float accumulations per pixel, round-half-even(sum / wsum).  No FMA: the products
// and sums are rounded one by one in the tap order, so the result equals the CPU loop bit for bit.
// One CTA = 32 x 8 outputs, one output per thread; the tile + halo, both LUTs and the per-row tap extents sit in
// shared memory (the colour LUT is read at data-dependent indices: shared memory, not constant memory).
constexpr int kBilTX = 32, kBilTY = 8;

struct BilateralLut {       // device copies, built once per context
    const float* color_w;   // [256]
    const float* space_w;   // [maxk] in tap order
    const int* jmax;        // [2 * radius + 1]: taps of row i are j in [-jmax, jmax]
    int radius, maxk;
};

inline size_t bilateral_smem(int radius, int maxk) {
    const int tw = kBilTX + 2 * radius, th = kBilTY + 2 * radius;
    size_t s = ((size_t)tw * th + 15) & ~(size_t)15;
    return s + (size_t)256 * 4 + (size_t)maxk * 4 + (size_t)(2 * radius + 1) * 4;
}

__global__ void __launch_bounds__(kBilTX * kBilTY) k_bilateral(const BlurJob* __restrict__ jobs, BilateralLut L, int rows,
                                                               int cols, int rx0, int ry0, int rx1, int ry1) {
    extern __shared__ __align__(16) unsigned char bsm[];
    const BlurJob job = jobs[blockIdx.z];
    const int r = L.radius, tw = kBilTX + 2 * r, th = kBilTY + 2 * r;
    unsigned char* tin = bsm;
    float* cw = reinterpret_cast<float*>(bsm + (((size_t)tw * th + 15) & ~(size_t)15));
    float* sw = cw + 256;
    int* jm = reinterpret_cast<int*>(sw + L.maxk);
    const int tid = threadIdx.y * kBilTX + threadIdx.x, nthr = kBilTX * kBilTY;
    const int bx = rx0 + blockIdx.x * kBilTX, by = ry0 + blockIdx.y * kBilTY;
    for (int i = tid; i < tw * th; i += nthr) {
        const int ly = i / tw, lx = i - ly * tw;
        tin[i] = job.src[(size_t)dev_reflect101(by - r + ly, rows) * job.sstep + dev_reflect101(bx - r + lx, cols)];
    }
    for (int i = tid; i < 256; i += nthr) cw[i] = L.color_w[i];
    for (int i = tid; i < L.maxk; i += nthr) sw[i] = L.space_w[i];
    for (int i = tid; i < 2 * r + 1; i += nthr) jm[i] = L.jmax[i];
    __syncthreads();
    const int gx = bx + threadIdx.x, gy = by + threadIdx.y;
    if (gx >= rx1 || gy >= ry1) return;
    const unsigned char* c0 = tin + (threadIdx.y + r) * tw + threadIdx.x + r;
    const int v0 = *c0;
    float sum = 0.f, wsum = 0.f;
    int k = 0;
    for (int i = -r; i <= r; ++i) {
        const int m = jm[i + r];
        const unsigned char* row = c0 + i * tw;
#pragma unroll 4
        for (int j = -m; j <= m; ++j, ++k) {
            const int v = row[j];
            const float w = __fmul_rn(sw[k], cw[abs(v - v0)]);
            sum = __fadd_rn(sum, __fmul_rn((float)v, w));
            wsum = __fadd_rn(wsum, w);
        }
    }
    job.dst[(size_t)gy * job.dstep + gx] = (unsigned char)__float2int_rn(__fdiv_rn(sum, wsum));
}

}  // namespace o3r
