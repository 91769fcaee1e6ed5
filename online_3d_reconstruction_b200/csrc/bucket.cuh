// bucket.cuh — the fused disparity -> cloud -> voxel engine (O3R_MERGE_ACCUMULATE_FUSED).
//
// What it replaces: createSingleImgPtCloud (pose_functions.cpp:1030-1134) + transformPtCloud (:1358-1362) +
// downsamplePtCloud(cloud, false) (:1654-1709, VoxelGrid leaf voxel_size / 5) + the append to cloud_big (pose.cpp:434)
// of a whole cycle, up to the partial sums the incremental merge (voxel.cuh engine 2) folds into the resident shard.
//
// Why not the radix sort of engine 1: PCL sorts the whole frame by leaf index only to GROUP the points of a leaf; a
// sort-based VoxelGrid moves ~110 B per point through HBM where 20 B are compulsory.  A leaf of voxel_size / 5 only holds
// points that are a few pixels apart, so the grouping is local: the cycle's points are binned ONCE into buckets of
// kBkB x kBkB leaf columns (all z) — nominally one combined-grid (voxel_size, voxel_size, 1000) cell — and each bucket
// (~80 points at 720p, voxel_size 0.05) is finished inside shared memory by one warp:
//
//   k_bk_hist     every pixel: validity, reprojection, transform -> leaf cell -> bucket; counts per bucket (RED), the
//                 frame's PCL bbox (for the int32 overflow guard)
//   k_bk_scan     exclusive scan of the bucket counts (decoupled look-back) + compact list of the non-empty buckets
//   k_bk_scatter  the same evaluation again + colour; the tile's points are staged in shared memory in scan order, every
//                 run of consecutive points of one bucket reserves its slots with ONE atomic, and the records leave as
//                 coalesced bursts: {x, y, z, rgb} (16 B) + scan position (4 B)
//   k_bk_reduce   one warp per bucket: counting sort by leaf column, rank inside the column by (z cell, scan position)
//                 -> the bucket in PCL's order restricted to it; every leaf is left-folded in scan order (bit-identical
//                 centroids: the oracle's stable order), each centroid is celled on the combined grid and the warp emits
//                 one partial cell per distinct combined cell (nearly always one), at an offset from a decoupled
//                 look-back so the partial list is reproducible.
//
// Determinism: the slot a point gets inside its bucket depends on an atomic race, which is why the scan position
// travels with it — the reduce orders by it, so every float is summed in a fixed order.  Leaf centroids are bit-exact;
// the combined-grid partial is a fixed tree (per lane in leaf order, then an xor butterfly), i.e. the same contract as
// O3R_MERGE_ACCUMULATE_TILED: keys, counts, colour sums exact, centroids equal to the oracle's left fold within float
// reassociation (<= 1e-5 relative), reproducible bit for bit from run to run.
//
// Limits (checked on the device; the host then reruns the batch through the sort engine and stays there):
// a bucket with more than kBkCapSlow points, more than kBkMaxPart distinct combined cells in one bucket, a point outside
// the host's conservative per-frame bucket grid.
#pragma once
#include "common.cuh"
#include "sort.cuh"
#include "stage_a.cuh"

namespace o3r {

constexpr int kBkB = 5;                            // bucket edge in leaf cells (= voxel_size / leaf)
constexpr int kBkSub = kBkB * kBkB;                // leaf columns per bucket (<= 32: one lane per column in the scan)
constexpr int kBkCap = 256;                        // points per bucket the reduce ranks in shared memory (fast path)
constexpr int kBkCapSlow = 65535;                  // larger buckets (parallax overlaps) are ranked one leaf column at a time; a column holds <= kBkCap
constexpr int kBkRounds = kBkCap / 32;
constexpr int kBkPerWarp = 4;                      // buckets a warp takes per ticket
constexpr int kBkMaxPart = 8;                      // partial cells per bucket
constexpr int kBkScanItems = 8;
constexpr int kBkScanTile = kThreads * kBkScanItems;
#ifndef O3R_BK_MINB
#define O3R_BK_MINB 3
#endif
constexpr int kBkKBits = 26;                       // (z cell - k0) must fit: (2^26 * 25) < 2^32

// per-frame bucket grid, from the host's conservative bound of the frame's world bbox
struct BkFrame {
    int i0c, j0c;        // leaf cell of the grid origin (multiples of kBkB)
    int ni, nj;          // extent in buckets
    int k0;              // lower bound of the z leaf cell
    uint32_t base;       // first bucket of the frame in the batch-wide arrays
    int pad0, pad1;
};

enum { BK_FLAG_RANGE = 1u, BK_FLAG_BUCKET = 2u, BK_FLAG_PART = 4u };

__device__ __forceinline__ void bk_cell(float x, float y, float z, float inv, int& i, int& j, int& k) {
    i = (int)floorf(__fmul_rn(x, inv));
    j = (int)floorf(__fmul_rn(y, inv));
    k = (int)floorf(__fmul_rn(z, inv));
}

// bucket of leaf column (i, j); `bad` is raised when the column lies outside the frame's grid (the index is then clamped:
// still a function of (i, j) only, and in bounds)
__device__ __forceinline__ uint32_t bk_index(const BkFrame& B, int i, int j, int k, bool& bad) {
    int bi = i - B.i0c, bj = j - B.j0c;
    const unsigned kr = (unsigned)(k - B.k0);
    if (bi < 0 || bj < 0 || kr >= (1u << kBkKBits)) { bad = true; bi = max(bi, 0); bj = max(bj, 0); }
    int I = bi / kBkB, J = bj / kBkB;
    if (I >= B.ni || J >= B.nj) { bad = true; I = min(I, B.ni - 1); J = min(J, B.nj - 1); }
    return B.base + (uint32_t)J * (uint32_t)B.ni + (uint32_t)I;
}

// ---- pass 1: bucket histogram + per-frame bbox ---------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(kThreads) k_bk_hist(AParams P, const FrameDev* __restrict__ frames,
                                                      const BkFrame* __restrict__ bk, float inv_f,
                                                      uint32_t* __restrict__ counts, uint32_t* __restrict__ bbox,
                                                      uint32_t* __restrict__ flags) {
    __shared__ double rl[256];
    __shared__ float zl[256];
    __shared__ uint32_t s_red[8];
    const int tile = blockIdx.x, f = blockIdx.y;
    const FrameDev& F = frames[f];
    const BkFrame B = bk[f];
    if (threadIdx.x < 8) s_red[threadIdx.x] = (threadIdx.x < 3) ? 0xffffffffu : 0u;
    load_lut(P, rl, zl);
    Samp S;
    eval4<DT>(P, F, rl, zl, tile, S);
    uint32_t mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    uint32_t cur = 0xffffffffu, run = 0;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (S.mask & (1u << j)) {
            float tx, ty, tz;
            xform(F.T, S.x[j], S.y[j], S.z[j], tx, ty, tz);
            const uint32_t ox = f2ord(tx), oy = f2ord(ty), oz = f2ord(tz);
            mn[0] = min(mn[0], ox); mx[0] = max(mx[0], ox);
            mn[1] = min(mn[1], oy); mx[1] = max(mx[1], oy);
            mn[2] = min(mn[2], oz); mx[2] = max(mx[2], oz);
            int ci, cj, ck;
            bk_cell(tx, ty, tz, inv_f, ci, cj, ck);
            const uint32_t b = bk_index(B, ci, cj, ck, bad);
            if (b == cur) ++run;
            else {
                if (run) atomicAdd(&counts[cur], run);
                cur = b; run = 1;
            }
        }
    if (run) atomicAdd(&counts[cur], run);
    if (bad) atomicOr(flags, BK_FLAG_RANGE);
    const uint32_t any = __ballot_sync(kFull, S.mask != 0u);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        mn[a] = __reduce_min_sync(kFull, mn[a]);
        mx[a] = __reduce_max_sync(kFull, mx[a]);
    }
    if ((threadIdx.x & 31) == 0 && any) {
        atomicAdd(&s_red[6], 1u);
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(&s_red[a], mn[a]); atomicMax(&s_red[3 + a], mx[a]); }
    }
    __syncthreads();
    if (threadIdx.x < 6 && s_red[6]) {
        if (threadIdx.x < 3) atomicMin(&bbox[f * 6 + threadIdx.x], s_red[threadIdx.x]);
        else atomicMax(&bbox[f * 6 + threadIdx.x], s_red[threadIdx.x]);
    }
}

// per frame: PCL's int32 overflow guard on the exact bbox (pass-through frames keep their points verbatim)
__global__ void k_bk_frames(int n_frames, const uint32_t* __restrict__ bbox, float inv_f, uint8_t* __restrict__ frame_pass,
                            uint32_t* __restrict__ frame_vox) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const GridParams G = make_grid(bbox + 6 * f, inv_f, inv_f, inv_f);
    frame_pass[f] = (uint8_t)(G.passthrough ? 1 : 0);
    frame_vox[f] = 0u;
}

// ---- exclusive scan of the bucket counts + list of the non-empty buckets -----------------------------------------------
// status word: flag << 62 | non-empty buckets << 34 | points
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr unsigned long long kBk64Local = 1ull << 62, kBk64Global = 2ull << 62, kBk64Mask = (1ull << 62) - 1ull;

// counts[g] becomes the bucket's first slot (the scatter's cursor); nl[r] = {first slot, points | frame << 16} of the
// r-th non-empty bucket; totals = {points, non-empty buckets}
__global__ void __launch_bounds__(kThreads) k_bk_scan(uint32_t* __restrict__ counts, uint32_t nb_total,
                                                      const BkFrame* __restrict__ bk, int n_frames, uint2* __restrict__ nl,
                                                      unsigned long long* __restrict__ status, uint32_t* __restrict__ ticket,
                                                      uint32_t* __restrict__ totals, uint32_t* __restrict__ flags) {
    __shared__ unsigned long long s_w[kWarps + 2];
    __shared__ uint32_t s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t t = s_ticket;
    const uint32_t nt = (nb_total + kBkScanTile - 1) / kBkScanTile;
    if (t >= nt) return;
    const uint32_t g0 = t * kBkScanTile + tid * kBkScanItems;
    uint32_t c[kBkScanItems];
    if (g0 + kBkScanItems <= nb_total) {
        const uint4 a = *reinterpret_cast<const uint4*>(counts + g0), b = *reinterpret_cast<const uint4*>(counts + g0 + 4);
        c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    } else {
#pragma unroll
        for (int q = 0; q < kBkScanItems; ++q) c[q] = g0 + q < nb_total ? counts[g0 + q] : 0u;
    }
    unsigned long long mine = 0;
    bool over = false;
#pragma unroll
    for (int q = 0; q < kBkScanItems; ++q) {
        mine += (unsigned long long)c[q] + (c[q] ? (1ull << 34) : 0ull);
        over = over || c[q] > (uint32_t)kBkCapSlow;
    }
    if (over) atomicOr(flags, BK_FLAG_BUCKET);
    // CTA-wide exclusive scan of the packed (non-empty, points) pair
    unsigned long long inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long w = lane < kWarps ? s_w[lane] : 0ull;
        unsigned long long winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(kFull, winc, o);
            if (lane >= o) winc += v;
        }
        if (lane < kWarps) s_w[lane] = winc - w;
        const unsigned long long tile_tot = __shfl_sync(kFull, winc, kWarps - 1);
        // decoupled look-back over the predecessors' totals, 32 tiles per step
        if (lane == 0) st_relaxed_u64(status + t, (t == 0 ? kBk64Global : kBk64Local) | tile_tot);
        unsigned long long pf = 0;
        for (int32_t back = (int32_t)t - 1; back >= 0; back -= 32) {
            const int32_t idx = back - lane;
            unsigned long long v = kBk64Global;
            if (idx >= 0)
                while (((v = ld_relaxed_u64(status + idx)) >> 62) == 0ull) __nanosleep(32);
            const unsigned gm = __ballot_sync(kFull, (v >> 62) == 2ull);
            const int first = gm ? __ffs(gm) - 1 : 31;
            unsigned long long part = lane <= first ? (v & kBk64Mask) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
            pf += part;
            if (gm) break;
        }
        if (lane == 0) {
            if (t > 0) st_relaxed_u64(status + t, kBk64Global | (pf + tile_tot));
            s_w[kWarps] = pf;
            if (t == nt - 1) {
                const unsigned long long all = pf + tile_tot;
                totals[0] = (uint32_t)(all & ((1ull << 34) - 1ull));
                totals[1] = (uint32_t)(all >> 34);
            }
        }
    }
    __syncthreads();
    unsigned long long ex = s_w[kWarps] + s_w[warp] + inc - mine;
    uint32_t start = (uint32_t)(ex & ((1ull << 34) - 1ull)), rank = (uint32_t)(ex >> 34);
    // frame of the thread's first bucket (the buckets of a frame are contiguous), advanced as g crosses a base
    int f = 0;
    if (g0 < nb_total) {
        int lo = 0, hi = n_frames - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (bk[mid].base <= g0) lo = mid; else hi = mid - 1;
        }
        f = lo;
    }
    uint32_t next_base = (f + 1 < n_frames) ? bk[f + 1].base : 0xffffffffu;
    uint32_t o[kBkScanItems];
#pragma unroll
    for (int q = 0; q < kBkScanItems; ++q) {
        const uint32_t g = g0 + q;
        o[q] = start;
        if (g < nb_total && c[q]) {
            while (g >= next_base) { ++f; next_base = (f + 1 < n_frames) ? bk[f + 1].base : 0xffffffffu; }
            nl[rank++] = make_uint2(start, min(c[q], 0xffffu) | ((uint32_t)f << 16));
            start += c[q];
        }
    }
    if (g0 + kBkScanItems <= nb_total) {
        *reinterpret_cast<uint4*>(counts + g0) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(counts + g0 + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int q = 0; q < kBkScanItems; ++q)
            if (g0 + q < nb_total) counts[g0 + q] = o[q];
    }
}

// Exclusive MAX scan of one value per thread over the CTA (identity 0).  `sm` is kWarps + 1 words; ends with a barrier.
__device__ __forceinline__ uint32_t block_excl_max_scan(uint32_t v, uint32_t* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc = max(inc, t);
    }
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    uint32_t before = 0;   // max over the warps in front of this one
#pragma unroll
    for (int w = 0; w < kWarps; ++w)
        if (w < warp) before = max(before, sm[w]);
    uint32_t prev = __shfl_up_sync(kFull, inc, 1);
    if (lane == 0) prev = 0;
    __syncthreads();
    return max(before, prev);
}

// ---- pass 2: evaluate again, colour, bin into the buckets ---------------------------------------------------------------
// pos_out = the point's position in the frame's scan order (keypoints, then the row-major grid): the order PCL's stable
// grouping (the oracle's) adds the points of a leaf in.
template <int DT>
__global__ void __launch_bounds__(kThreads) k_bk_scatter(AParams P, const FrameDev* __restrict__ frames,
                                                         const BkFrame* __restrict__ bk, float inv_f,
                                                         uint32_t* __restrict__ cursor, float4* __restrict__ pts,
                                                         uint32_t* __restrict__ pos_out, const uint32_t* __restrict__ flags) {
    __shared__ double rl[256];
    __shared__ float zl[256];
    __shared__ uint32_t s_scan[34];
    __shared__ __align__(16) float4 s_pts[kTileA];
    __shared__ uint32_t s_pos[kTileA];
    __shared__ __align__(16) uint32_t s_bkt[kTileA + 4];   // bucket of the staged point, then its destination slot
    __shared__ uint32_t s_base[kTileA];                    // by run head position: first slot of the run
    if (*flags) return;   // the histogram already overflowed: the host reruns the batch through the sort engine
    const int tile = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const FrameDev& F = frames[f];
    const BkFrame B = bk[f];
    load_lut(P, rl, zl);
    Samp S;
    eval4<DT>(P, F, rl, zl, tile, S);
    uint32_t total;
    const uint32_t off = block_excl_scan((uint32_t)__popc(S.mask), s_scan, total);
    if (total == 0) return;
    if (S.mask) {
        uint32_t rgb[4];
        if (S.vec_row) {
            const uint32_t* c = reinterpret_cast<const uint32_t*>(F.bgr + (size_t)S.py[0] * F.bgr_step + 3 * (size_t)S.px[0]);
            const uint32_t w0 = __ldcs(c), w1 = __ldcs(c + 1), w2 = __ldcs(c + 2);
            rgb[0] = w0 & 0x00ffffffu;
            rgb[1] = (w0 >> 24) | ((w1 & 0xffffu) << 8);
            rgb[2] = (w1 >> 16) | ((w2 & 0xffu) << 16);
            rgb[3] = w2 >> 8;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (S.mask & (1u << j)) {
                    const uint8_t* c = F.bgr + (size_t)S.py[j] * F.bgr_step + 3 * (size_t)S.px[j];
                    rgb[j] = ((uint32_t)c[2] << 16) | ((uint32_t)c[1] << 8) | (uint32_t)c[0];
                }
        }
        // scan position: keypoint index, or n_kp + grid sample index
        const uint32_t p0 = tile < P.kp_tiles ? (uint32_t)tile * kTileA + tid * 4
                                              : (uint32_t)F.n_kp + (uint32_t)(tile - P.kp_tiles) * kTileA + tid * 4;
        uint32_t o = off;
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (S.mask & (1u << j)) {
                float tx, ty, tz;
                xform(F.T, S.x[j], S.y[j], S.z[j], tx, ty, tz);
                int ci, cj, ck;
                bk_cell(tx, ty, tz, inv_f, ci, cj, ck);
                s_pts[o] = make_float4(tx, ty, tz, __uint_as_float(rgb[j]));
                s_pos[o] = p0 + j;
                s_bkt[o] = bk_index(B, ci, cj, ck, bad);
                ++o;
            }
    }
    __syncthreads();
    // ---- runs of consecutive staged points of one bucket: the run's last point reserves the slots
    const uint32_t e0 = tid * 4;
    uint32_t b[6];   // b[1 + j] = bucket of staged point e0 + j; b[0] / b[5] = the neighbours
#pragma unroll
    for (int j = -1; j <= 4; ++j) {
        const uint32_t e = e0 + j;   // (wraps for e0 == 0, j == -1: caught by e < total being false for 0xffffffff)
        b[j + 1] = e < total ? s_bkt[e] : 0xffffffffu;
    }
    uint32_t lh = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (e0 + j < total && b[j + 1] != b[j]) lh = e0 + j;
    uint32_t L = block_excl_max_scan(lh, s_scan);   // head of the run that is open when this thread's points begin
    uint32_t hp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t e = e0 + j;
        hp[j] = 0;
        if (e < total) {
            if (b[j + 1] != b[j]) L = e;
            hp[j] = L;
            if (b[j + 2] != b[j + 1]) s_base[L] = atomicAdd(&cursor[b[j + 1]], e - L + 1u);
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t e = e0 + j;
        if (e < total) s_bkt[e] = s_base[hp[j]] + (e - hp[j]);
    }
    __syncthreads();
    for (uint32_t i = tid; i < total; i += kThreads) {
        const uint32_t g = s_bkt[i];
        pts[g] = s_pts[i];
        pos_out[g] = s_pos[i];
    }
}

// ---- pass 3: one warp per bucket ---------------------------------------------------------------------------------------
// Warps are independent (no CTA barrier, no ordering between buckets): the r-th non-empty bucket writes its main partial cell
// to slot r of the chunk's partial list; the rare stray centroids (see below) are appended behind the main slots by an
// atomic counter and put into a canonical order afterwards (k_bk_strays), so the list is reproducible.
struct BkWarp {
    float4 pts[kBkCap];                       // the points being ranked, load order
    union {
        unsigned long long bkeys[kBkCap];     // (leaf key << 32 | scan position) in column order
        struct { uint32_t key[kBkCap]; uint8_t idx[kBkCap]; } s;   // the same in final order: leaf key, index into pts
        uint32_t ppos[kBkCap];                // big buckets: scan positions of the staged column (dead once the keys are formed)
    } u;
    o3r_cell stray[kBkMaxPart];               // centroids that fell into a neighbour of the nominal combined cell
    uint16_t bin[36];                         // per column: count, then first position
};
constexpr size_t bk_reduce_smem() { return sizeof(BkWarp) * kWarps; }

// ordering key of a point inside its bucket: leaf (z cell, column) — or the column alone in a pass-through frame, where
// every point is its own voxel — then the scan position
__device__ __forceinline__ unsigned long long bk_key64(const BkFrame& B, bool pass, float inv_f, const float4& p, uint32_t ps,
                                                       uint32_t& col) {
    int ci, cj, ck;
    bk_cell(p.x, p.y, p.z, inv_f, ci, cj, ck);
    const int bi = ci - B.i0c, bj = cj - B.j0c;
    col = (uint32_t)(bi - (bi / kBkB) * kBkB) + (uint32_t)(bj - (bj / kBkB) * kBkB) * kBkB;
    const uint32_t k32 = pass ? col : (uint32_t)(ck - B.k0) * kBkSub + col;
    return ((unsigned long long)k32 << 32) | ps;
}

// V1 centroid of pcl::CentroidPoint<PointXYZRGB> (float sums / (float)n; colour uint32(sum / n))
__device__ __forceinline__ float4 bk_centroid(float sx, float sy, float sz, uint32_t n, uint32_t r, uint32_t g, uint32_t b) {
    const float fn = (float)n;
    const uint32_t rgb = ((uint32_t)__fdiv_rn((float)r, fn) << 16) | ((uint32_t)__fdiv_rn((float)g, fn) << 8) |
                         (uint32_t)__fdiv_rn((float)b, fn);
    return make_float4(__fdiv_rn(sx, fn), __fdiv_rn(sy, fn), __fdiv_rn(sz, fn), __uint_as_float(rgb));
}

// running state of one bucket across the passes of bk_pass
struct BkAcc {
    unsigned long long K0;    // nominal combined cell (x, y fields) + the first centroid's z cell
    float ax, ay, az;         // per lane: the lane's centroids of cell K0, added in leaf order
    uint32_t an, ar, ag, ab;
    uint32_t nvox, nstray;    // warp-uniform
    int have_k0;
};

// Ranks the m (<= kBkCap) points src_pts[0..m) / src_pos[0..m) — a whole bucket, or one leaf column of a big bucket —
// into PCL's order (column, z cell, scan position), left-folds every leaf in scan order, cells each centroid on the
// combined grid and adds it to the bucket's partial sums.  STAGED: the points already sit in W.pts (src_pts == W.pts).
template <bool STAGED>
__device__ __forceinline__ void bk_pass(BkWarp& W, const float4* __restrict__ src_pts, const uint32_t* __restrict__ src_pos,
                                        uint32_t m, const BkFrame& B, bool pass, float inv_f, float icx, float icz, BkAcc& A,
                                        int* cmn, int* cmx, float4* __restrict__ dbg_vox, uint32_t* __restrict__ dbg_cnt) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int R = (int)((m + 31u) >> 5);
    W.bin[lane] = 0;
    if (lane < 4) W.bin[32 + lane] = 0;
    __syncwarp();
    // ---- load; counting sort by leaf column (stable in load order, which is arbitrary — the rank below orders)
    unsigned long long key64[kBkRounds];
    uint32_t sb[kBkRounds];   // column | slot inside the column << 8, later the final position
#pragma unroll
    for (int rr = 0; rr < kBkRounds; ++rr) {
        if (rr >= R) break;
        const uint32_t e = rr * 32 + lane;
        const bool valid = e < m;
        const unsigned vm = __ballot_sync(kFull, valid);
        if (valid) {
            float4 p;
            uint32_t ps;
            if (STAGED) { p = W.pts[e]; ps = W.u.ppos[e]; }
            else { p = __ldcs(src_pts + e); ps = __ldcs(src_pos + e); W.pts[e] = p; }
            uint32_t s;
            key64[rr] = bk_key64(B, pass, inv_f, p, ps, s);
            const unsigned peers = __match_any_sync(vm, s);
            const uint32_t old = W.bin[s];
            __syncwarp(vm);
            if ((peers & lt) == 0u) W.bin[s] = (uint16_t)(old + __popc(peers));
            __syncwarp(vm);
            sb[rr] = s | ((old + __popc(peers & lt)) << 8);
        }
    }
    __syncwarp();
    {   // first position of every column
        const uint32_t cv = W.bin[lane];
        uint32_t inc = cv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += v;
        }
        __syncwarp();
        W.bin[lane] = (uint16_t)(inc - cv);
        if (lane == 31) W.bin[32] = (uint16_t)inc;
    }
    __syncwarp();   // (STAGED: every lane has read its scan positions, which the keys now overwrite)
#pragma unroll
    for (int rr = 0; rr < kBkRounds; ++rr) {
        if (rr >= R) break;
        if ((uint32_t)(rr * 32 + lane) < m) W.u.bkeys[(uint32_t)W.bin[sb[rr] & 255u] + (sb[rr] >> 8)] = key64[rr];
    }
    __syncwarp();
    // ---- rank inside the column by (leaf key, scan position): final position = PCL's order inside the bucket
#pragma unroll
    for (int rr = 0; rr < kBkRounds; ++rr) {
        if (rr >= R) break;
        if ((uint32_t)(rr * 32 + lane) < m) {
            const uint32_t s = sb[rr] & 255u;
            const uint32_t a = W.bin[s], e = W.bin[s + 1];
            uint32_t rk = 0;
            for (uint32_t i = a; i < e; ++i) rk += W.u.bkeys[i] < key64[rr] ? 1u : 0u;
            sb[rr] = a + rk;
        }
    }
    __syncwarp();
#pragma unroll
    for (int rr = 0; rr < kBkRounds; ++rr) {
        if (rr >= R) break;
        if ((uint32_t)(rr * 32 + lane) < m) {
            W.u.s.key[sb[rr]] = (uint32_t)(key64[rr] >> 32);
            W.u.s.idx[sb[rr]] = (uint8_t)(rr * 32 + lane);
        }
    }
    __syncwarp();
    // ---- leaves: left fold in scan order; combined-grid cell of every centroid; the bucket's partial sums
#pragma unroll 1
    for (int rr = 0; rr < R; ++rr) {
        const uint32_t fp = rr * 32 + lane;
        bool head = false;
        uint32_t k32 = 0;
        if (fp < m) {
            k32 = W.u.s.key[fp];
            head = pass || fp == 0 || W.u.s.key[fp - 1] != k32;
        }
        float4 cen = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned long long vk = 0;
        if (head) {
            float4 p = W.pts[W.u.s.idx[fp]];
            if (pass) {
                cen = p;   // PCL: output = *input_
            } else {
                float sx = 0.f, sy = 0.f, sz = 0.f;
                uint32_t n = 0, cr = 0, cg = 0, cb = 0;
                uint32_t q = fp;
                for (;;) {
                    sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z);
                    const uint32_t w = __float_as_uint(p.w);
                    cr += (w >> 16) & 255u; cg += (w >> 8) & 255u; cb += w & 255u;
                    ++n; ++q;
                    if (q >= m || W.u.s.key[q] != k32) break;
                    p = W.pts[W.u.s.idx[q]];
                }
                cen = bk_centroid(sx, sy, sz, n, cr, cg, cb);
            }
            if (dbg_vox) dbg_vox[atomicAdd(dbg_cnt, 1u)] = cen;
            cen.z = __fadd_rn(cen.z, 500.0f);   // pose_functions.cpp:1666
            const int vi = (int)floorf(__fmul_rn(cen.x, icx)), vj = (int)floorf(__fmul_rn(cen.y, icx)),
                      vkz = (int)floorf(__fmul_rn(cen.z, icz));
            cmn[0] = min(cmn[0], vi); cmx[0] = max(cmx[0], vi);
            cmn[1] = min(cmn[1], vj); cmx[1] = max(cmx[1], vj);
            cmn[2] = min(cmn[2], vkz); cmx[2] = max(cmx[2], vkz);
            const long long Bi = 1 << 20;
            vk = ((unsigned long long)(vkz + Bi) << 42) | ((unsigned long long)(vj + Bi) << 21) | (unsigned long long)(vi + Bi);
        }
        if (!A.have_k0) {   // (warp-uniform) the first centroid of the bucket fixes the z cell of the nominal key
            const unsigned long long first = __shfl_sync(kFull, vk, 0);   // position 0 is always a head
            A.K0 = (first & ~((1ull << 42) - 1ull)) | A.K0;
            A.have_k0 = 1;
        }
        const unsigned hm = __ballot_sync(kFull, head);
        A.nvox += __popc(hm);
        const bool other = head && vk != A.K0;
        if (head && !other) {
            const uint32_t w = __float_as_uint(cen.w);
            A.ax = __fadd_rn(A.ax, cen.x); A.ay = __fadd_rn(A.ay, cen.y); A.az = __fadd_rn(A.az, cen.z);
            ++A.an; A.ar += (w >> 16) & 255u; A.ag += (w >> 8) & 255u; A.ab += w & 255u;
        }
        const unsigned om = __ballot_sync(kFull, other);
        if (om) {   // a centroid on the border of the nominal cell to the last float bit: its own record
            if (other) {
                const uint32_t sl = A.nstray + __popc(om & lt);
                if (sl < (uint32_t)kBkMaxPart) {
                    const uint32_t w = __float_as_uint(cen.w);
                    o3r_cell pc;
                    pc.key = vk; pc.sx = cen.x; pc.sy = cen.y; pc.sz = cen.z; pc.n = 1u;
                    pc.sr = (w >> 16) & 255u; pc.sg = (w >> 8) & 255u; pc.sb = w & 255u; pc.pad = 0u;
                    W.stray[sl] = pc;
                }
            }
            A.nstray += __popc(om);
        }
    }
    __syncwarp();
}

// out[*out_base + r]            main partial of the r-th non-empty bucket (r < totals[1])
// out[*out_base + totals[1] + ...]  strays, appended through *stray_cnt (at most stray_cap)
__global__ void __launch_bounds__(kThreads, O3R_BK_MINB) k_bk_reduce(
    const float4* __restrict__ pts, const uint32_t* __restrict__ pos, const uint2* __restrict__ nl,
    const uint32_t* __restrict__ totals, const BkFrame* __restrict__ bk, const uint8_t* __restrict__ frame_pass, float inv_f,
    float icx, float icz, o3r_cell* __restrict__ out, const uint32_t* __restrict__ out_base, uint32_t* __restrict__ stray_cnt,
    uint32_t stray_cap, uint32_t* __restrict__ ticket, uint32_t* __restrict__ frame_vox, int* __restrict__ cellbb,
    uint32_t* __restrict__ flags, float4* __restrict__ dbg_vox, uint32_t* __restrict__ dbg_cnt) {
    extern __shared__ __align__(16) unsigned char bk_smem_raw[];
    __shared__ int s_bb[6];
    if (flags[0]) return;   // raised by the histogram / scan (complete before this kernel starts): uniform exit
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    BkWarp& W = reinterpret_cast<BkWarp*>(bk_smem_raw)[warp];
    const uint32_t NR = totals[1];
    o3r_cell* const out0 = out + *out_base;
    if (tid < 6) s_bb[tid] = tid < 3 ? 0x7fffffff : (int)0x80000000;
    int cmn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, cmx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    for (;;) {
        uint32_t r0 = 0;
        if (lane == 0) r0 = atomicAdd(ticket, (uint32_t)kBkPerWarp);
        r0 = __shfl_sync(kFull, r0, 0);
        if (r0 >= NR) break;
#pragma unroll 1
        for (uint32_t r = r0; r < min(r0 + (uint32_t)kBkPerWarp, NR); ++r) {
            const uint2 ent = nl[r];
            const uint32_t start = ent.x, cnt = ent.y & 0xffffu;
            const int f = (int)(ent.y >> 16);
            const BkFrame B = bk[f];
            const bool pass = frame_pass[f] != 0;
            const float4* gp = pts + start;
            const uint32_t* gpos = pos + start;
            BkAcc A;
            A.ax = A.ay = A.az = 0.f;
            A.an = A.ar = A.ag = A.ab = 0u;
            A.nvox = A.nstray = 0u;
            A.have_k0 = 0;
            {   // the bucket's NOMINAL combined cell: its leaf columns are [5I, 5I + 5) x [5J, 5J + 5), i.e. cell (I, J) — a
                // centroid lands in a neighbour only when it sits on the cell border to the last float bit
                const float4 p = gp[0];
                int ci, cj, ck;
                bk_cell(p.x, p.y, p.z, inv_f, ci, cj, ck);
                const long long Bi = 1 << 20;
                const long long nI = ci >= 0 ? ci / kBkB : -((-ci + kBkB - 1) / kBkB), nJ = cj >= 0 ? cj / kBkB : -((-cj + kBkB - 1) / kBkB);
                A.K0 = ((unsigned long long)(nJ + Bi) << 21) | (unsigned long long)(nI + Bi);
            }
            if (cnt <= (uint32_t)kBkCap) {
                bk_pass<false>(W, gp, gpos, cnt, B, pass, inv_f, icx, icz, A, cmn, cmx, dbg_vox, dbg_cnt);
            } else {
                // ---- rare (overlapping surfaces pile up in one cell): one leaf column at a time.  PCL's order is
                // (column, z cell, scan position), so the columns can be ranked and folded one after the other.
                for (uint32_t col = 0; col < (uint32_t)kBkSub; ++col) {
                    uint32_t m = 0;
                    for (uint32_t e0 = 0; e0 < cnt; e0 += 32) {
                        const uint32_t e = e0 + lane;
                        bool mine = false;
                        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                        uint32_t ps = 0;
                        if (e < cnt) {
                            p = gp[e]; ps = gpos[e];
                            uint32_t c;
                            bk_key64(B, pass, inv_f, p, ps, c);
                            mine = c == col;
                        }
                        const unsigned mm = __ballot_sync(kFull, mine);
                        if (mine) {
                            const uint32_t at = m + __popc(mm & ((1u << lane) - 1u));
                            if (at < (uint32_t)kBkCap) { W.pts[at] = p; W.u.ppos[at] = ps; }
                        }
                        m += __popc(mm);
                    }
                    __syncwarp();
                    if (m > (uint32_t)kBkCap) {   // more than 256 points in ONE leaf column: give up (host falls back)
                        if (lane == 0) atomicOr(flags + 1, BK_FLAG_BUCKET);
                        m = kBkCap;
                    }
                    if (m) bk_pass<true>(W, W.pts, W.u.ppos, m, B, pass, inv_f, icx, icz, A, cmn, cmx, dbg_vox, dbg_cnt);
                }
            }
            // the main partial: lanes already folded their leaves in order; fixed xor butterfly across the lanes
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                A.ax = __fadd_rn(A.ax, __shfl_xor_sync(kFull, A.ax, o));
                A.ay = __fadd_rn(A.ay, __shfl_xor_sync(kFull, A.ay, o));
                A.az = __fadd_rn(A.az, __shfl_xor_sync(kFull, A.az, o));
            }
            const uint32_t an = __reduce_add_sync(kFull, A.an), ar = __reduce_add_sync(kFull, A.ar),
                           ag = __reduce_add_sync(kFull, A.ag), ab = __reduce_add_sync(kFull, A.ab);
            uint32_t ns = A.nstray;
            if (ns > (uint32_t)kBkMaxPart) {
                if (lane == 0) atomicOr(flags + 1, BK_FLAG_PART);
                ns = kBkMaxPart;
            }
            o3r_cell mainc;
            mainc.key = A.K0; mainc.sx = A.ax; mainc.sy = A.ay; mainc.sz = A.az; mainc.n = an;
            mainc.sr = ar; mainc.sg = ag; mainc.sb = ab; mainc.pad = 0u;
            if (an == 0u) mainc = W.stray[--ns];   // every centroid strayed: no empty record, the last stray takes the main slot
            if (lane == 0) {
                out0[r] = mainc;
                atomicAdd(&frame_vox[f], A.nvox);
            }
            if (ns) {
                uint32_t at = 0;
                if (lane == 0) at = atomicAdd(stray_cnt, ns);
                at = __shfl_sync(kFull, at, 0);
                if (at + ns > stray_cap) {
                    if (lane == 0) atomicOr(flags + 1, BK_FLAG_PART);
                } else {
                    uint32_t* dst = reinterpret_cast<uint32_t*>(out0 + NR + at);
                    const uint32_t* src = reinterpret_cast<const uint32_t*>(&W.stray[0]);
                    for (uint32_t i = lane; i < ns * (uint32_t)(sizeof(o3r_cell) / 4); i += 32) dst[i] = src[i];
                }
            }
            __syncwarp();
        }
    }
    // ---- range of combined-grid cells touched (the merge packs its sort keys into it)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int lo = __reduce_min_sync(kFull, cmn[a]), hi = __reduce_max_sync(kFull, cmx[a]);
        cmn[a] = lo; cmx[a] = hi;
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(&s_bb[a], cmn[a]); atomicMax(&s_bb[3 + a], cmx[a]); }
    }
    __syncthreads();
    if (tid < 3) { if (s_bb[tid] != 0x7fffffff) atomicMin(&cellbb[tid], s_bb[tid]); }
    else if (tid < 6) { if (s_bb[tid] != (int)0x80000000) atomicMax(&cellbb[tid], s_bb[tid]); }
}

// Canonical order of the chunk's stray records (they were appended in atomic order): rank by the whole record.  Two records
// that compare equal are identical, so their mutual order does not matter.  One CTA; n is tiny (a couple per frame).
// Also publishes the chunk's record count: main slots + strays.
__device__ __forceinline__ bool bk_cell_less(const o3r_cell& a, const o3r_cell& b) {
    if (a.key != b.key) return a.key < b.key;
    const uint32_t av[7] = {__float_as_uint(a.sx), __float_as_uint(a.sy), __float_as_uint(a.sz), a.n, a.sr, a.sg, a.sb};
    const uint32_t bv[7] = {__float_as_uint(b.sx), __float_as_uint(b.sy), __float_as_uint(b.sz), b.n, b.sr, b.sg, b.sb};
#pragma unroll
    for (int i = 0; i < 7; ++i)
        if (av[i] != bv[i]) return av[i] < bv[i];
    return false;
}
__global__ void __launch_bounds__(1024) k_bk_strays(o3r_cell* __restrict__ out, const uint32_t* __restrict__ out_base,
                                                    const uint32_t* __restrict__ totals, const uint32_t* __restrict__ stray_cnt,
                                                    uint32_t stray_cap, o3r_cell* __restrict__ tmp, uint32_t* __restrict__ chunk_total,
                                                    const uint32_t* __restrict__ flags) {
    if (flags[0]) { if (threadIdx.x == 0) *chunk_total = 0u; return; }
    const uint32_t NR = totals[1], n = min(*stray_cnt, stray_cap);
    if (threadIdx.x == 0) *chunk_total = NR + n;
    if (n < 2) return;
    o3r_cell* s = out + *out_base + NR;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const o3r_cell me = s[i];
        uint32_t rk = 0;
        for (uint32_t j = 0; j < n; ++j) {
            const o3r_cell o = s[j];
            rk += (bk_cell_less(o, me) || (!bk_cell_less(me, o) && j < i)) ? 1u : 0u;
        }
        tmp[rk] = me;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s[i] = tmp[i];
}

}  // namespace o3r
