// voxel.cuh — pcl::VoxelGrid-equivalent downsample on sorted (leaf index, point) pairs.
//
// Engine 1 ("grid"): PCL's own relative 32-bit leaf index per segment (frame), 4-pass radix sort,
//   run detection and a sequential in-order sum per run (CentroidPoint semantics).  Used for the
//   per-frame grid (leaf voxel_size/5), for the one-shot combined grid of RETAIN mode and for the
//   stand-alone o3r_voxel_grid probe.  Because the sort is stable and one thread sums a run in order,
//   centroids are bit-identical to the oracle's, not merely within tolerance.
// Engine 2 ("acc"): the incremental global merge.  64-bit absolute cell keys on the combined grid
//   (leaf v, v, 1000 after z += 500), sort of the cycle's points, and per run a sequential sum that
//   STARTS from the resident accumulator of that cell, so the resident sums equal what a one-shot
//   voxelisation of the whole history would have produced (pose.cpp:434 + :527-531).
#pragma once
#include "common.cuh"
#include "sort.cuh"

namespace o3r {

constexpr int kTileV = kThreads * 4;  // 1024 sorted elements per CTA in the run kernels

// ---- bbox of arbitrary point segments (PCL getMinMax3D), z optionally shifted by +500 ------------------------
__global__ void __launch_bounds__(kThreads) k_bbox_init(uint32_t* bbox, int n_seg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_seg * 6) bbox[i] = (i % 6 < 3) ? 0xffffffffu : 0u;
}

__global__ void __launch_bounds__(kThreads) k_bbox_pts(const float4* __restrict__ pts, const uint32_t* __restrict__ seg_off,
                                                       int z_shift, uint32_t* __restrict__ bbox) {
    __shared__ uint32_t s_red[6];
    const int s = blockIdx.y;
    const uint32_t beg = seg_off[s], n = seg_off[s + 1] - beg;
    if (blockIdx.x * kTileV >= n) return;
    if (threadIdx.x < 6) s_red[threadIdx.x] = (threadIdx.x < 3) ? 0xffffffffu : 0u;
    __syncthreads();
    uint32_t mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t k = blockIdx.x * kTileV + j * kThreads + threadIdx.x;
        if (k < n) {
            const float4 p = pts[beg + k];
            const float z = z_shift ? __fadd_rn(p.z, 500.0f) : p.z;
            const uint32_t ox = f2ord(p.x), oy = f2ord(p.y), oz = f2ord(z);
            mn[0] = min(mn[0], ox); mx[0] = max(mx[0], ox);
            mn[1] = min(mn[1], oy); mx[1] = max(mx[1], oy);
            mn[2] = min(mn[2], oz); mx[2] = max(mx[2], oz);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        mn[a] = __reduce_min_sync(kFull, mn[a]);
        mx[a] = __reduce_max_sync(kFull, mx[a]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(&s_red[a], mn[a]); atomicMax(&s_red[3 + a], mx[a]); }
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(&bbox[s * 6 + threadIdx.x], s_red[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(&bbox[s * 6 + threadIdx.x], s_red[threadIdx.x]);
}

__global__ void k_grid_params(int n_seg, const uint32_t* __restrict__ bbox, float ix, float iy, float iz,
                              GridParams* __restrict__ grids) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_seg) grids[s] = make_grid(bbox + 6 * s, ix, iy, iz);
}

// leaf index of arbitrary points (engine 1 when the points do not come from k_emit)
__global__ void __launch_bounds__(kThreads) k_vg_key(const float4* __restrict__ pts, const uint32_t* __restrict__ seg_off,
                                                     const GridParams* __restrict__ grids, int z_shift,
                                                     uint32_t* __restrict__ keys) {
    const int s = blockIdx.y;
    const uint32_t beg = seg_off[s], n = seg_off[s + 1] - beg;
    const GridParams G = grids[s];
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const float4 p = pts[beg + i];
        const float z = z_shift ? __fadd_rn(p.z, 500.0f) : p.z;
        keys[beg + i] = G.passthrough ? i : vg_rel_idx(G, p.x, p.y, z);
    }
}

// ---- engine 1: run heads -> counts; reduce -> centroids ----------------------------------------------------------
struct VgArgs {
    const uint32_t* keys0; const uint32_t* keys1;
    const uint32_t* vals0; const uint32_t* vals1;
    const uint32_t* seg_off;
    const SortPlan* plan;
    const GridParams* grids;     // may be null (no passthrough handling needed beyond keys)
    const float4* pts;
    uint32_t tiles_ub;           // per segment, of kTileV
    uint32_t min_points;
    int z_shift;                 // sums use z + 500 and outputs subtract it again (combined grid)
    float lx_inv, ly_inv, lz_inv;  // for the optional absolute-key output
};

__device__ __forceinline__ bool vg_emits(const uint32_t* keys, uint32_t pos, uint32_t n, uint32_t min_points) {
    if (pos > 0 && keys[pos] == keys[pos - 1]) return false;
    if (min_points <= 1) return true;
    const uint32_t k = keys[pos];
    uint32_t len = 1;
    while (len < min_points && pos + len < n && keys[pos + len] == k) ++len;
    return len >= min_points;
}

__global__ void __launch_bounds__(kThreads) k_vg_heads(VgArgs A, uint32_t* __restrict__ head_cnt) {
    const int s = blockIdx.y;
    const uint32_t t = blockIdx.x;
    const uint32_t beg = A.seg_off[s], n = A.seg_off[s + 1] - beg;
    uint32_t c = 0;
    if (t * kTileV < n) {
        const uint32_t* keys = (A.plan[s].final_parity ? A.keys1 : A.keys0) + beg;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t pos = t * kTileV + threadIdx.x * 4 + j;
            if (pos < n && vg_emits(keys, pos, n, A.min_points)) ++c;
        }
    }
    c = __reduce_add_sync(kFull, c);
    __shared__ uint32_t s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_c, c);
    __syncthreads();
    if (threadIdx.x == 0) head_cnt[(size_t)s * A.tiles_ub + t] = s_c;
}

__global__ void __launch_bounds__(kThreads) k_vg_reduce(VgArgs A, const uint32_t* __restrict__ head_off,
                                                        const uint32_t* __restrict__ head_total,
                                                        float4* __restrict__ out, uint32_t* __restrict__ seg_out_off,
                                                        int n_seg, uint64_t* __restrict__ out_keys,
                                                        uint32_t* __restrict__ out_counts) {
    __shared__ uint32_t s_scan[34];
    const int s = blockIdx.y;
    const uint32_t t = blockIdx.x;
    if (t == 0 && threadIdx.x == 0) {
        seg_out_off[s] = head_off[(size_t)s * A.tiles_ub];
        if (s == n_seg - 1) seg_out_off[n_seg] = *head_total;
    }
    const uint32_t beg = A.seg_off[s], n = A.seg_off[s + 1] - beg;
    if (t * kTileV >= n) return;
    const int par = A.plan[s].final_parity;
    const uint32_t* keys = (par ? A.keys1 : A.keys0) + beg;
    const uint32_t* vals = (par ? A.vals1 : A.vals0) + beg;
    uint32_t flags = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t pos = t * kTileV + threadIdx.x * 4 + j;
        if (pos < n && vg_emits(keys, pos, n, A.min_points)) flags |= 1u << j;
    }
    uint32_t tot;
    uint32_t o = head_off[(size_t)s * A.tiles_ub + t] + block_excl_scan((uint32_t)__popc(flags), s_scan, tot);
    if (!flags) return;
    const bool verbatim = A.grids && A.grids[s].passthrough;  // PCL: output = *input_
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (!(flags & (1u << j))) continue;
        const uint32_t pos = t * kTileV + threadIdx.x * 4 + j;
        const uint32_t k = keys[pos];
        if (verbatim) {
            const float4 p = A.pts[vals[pos]];
            out[o] = p;
            if (out_keys) out_keys[o] = abs_cell_key(p.x, p.y, A.z_shift ? __fadd_rn(p.z, 500.0f) : p.z, A.lx_inv, A.ly_inv, A.lz_inv);
            if (out_counts) out_counts[o] = 1;
            ++o;
            continue;
        }
        // CentroidPoint<PointXYZRGB>: float sums in run order, divide by (float)n; colour sums as floats
        float sx = 0.f, sy = 0.f, sz = 0.f, sr = 0.f, sg = 0.f, sb = 0.f;
        uint32_t cnt = 0;
        float4 first = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t q = pos; q < n && keys[q] == k; ++q) {
            const float4 p = A.pts[vals[q]];
            const float z = A.z_shift ? __fadd_rn(p.z, 500.0f) : p.z;
            if (cnt == 0) first = make_float4(p.x, p.y, z, 0.f);
            const uint32_t c = __float_as_uint(p.w);
            sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, z);
            sr = __fadd_rn(sr, (float)((c >> 16) & 255u));
            sg = __fadd_rn(sg, (float)((c >> 8) & 255u));
            sb = __fadd_rn(sb, (float)(c & 255u));
            ++cnt;
        }
        const float fn = (float)cnt;
        float cz = __fdiv_rn(sz, fn);
        if (A.z_shift) cz = __fsub_rn(cz, 500.0f);
        const uint32_t rgb = ((uint32_t)__fdiv_rn(sr, fn) << 16) | ((uint32_t)__fdiv_rn(sg, fn) << 8) |
                             (uint32_t)__fdiv_rn(sb, fn);
        out[o] = make_float4(__fdiv_rn(sx, fn), __fdiv_rn(sy, fn), cz, __uint_as_float(rgb));
        if (out_keys) out_keys[o] = abs_cell_key(first.x, first.y, first.z, A.lx_inv, A.ly_inv, A.lz_inv);
        if (out_counts) out_counts[o] = cnt;
        ++o;
    }
}

// passthrough segments keep their points verbatim (PCL: output = *input_); the identity keys already make
// every element its own run, so k_vg_reduce reproduces them — except that a centroid of one point is
// sum/1 = the point and uint32(c/1) = c, i.e. bit-identical.  Nothing extra to do.

// ---- engine 2: incremental merge on the combined grid -----------------------------------------------------------
struct AccItemsPts {   // items are points of weight 1 (z gets +500)
    const float4* pts;
    __device__ __forceinline__ void add(uint32_t v, float& sx, float& sy, float& sz, uint32_t& n, uint32_t& sr,
                                        uint32_t& sg, uint32_t& sb) const {
        const float4 p = pts[v];
        const uint32_t c = __float_as_uint(p.w);
        sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, __fadd_rn(p.z, 500.0f));
        n += 1; sr += (c >> 16) & 255u; sg += (c >> 8) & 255u; sb += c & 255u;
    }
};
struct AccItemsCells {  // items are partial cells received from other ranks
    const o3r_cell* cells;
    __device__ __forceinline__ void add(uint32_t v, float& sx, float& sy, float& sz, uint32_t& n, uint32_t& sr,
                                        uint32_t& sg, uint32_t& sb) const {
        const o3r_cell c = cells[v];
        sx = __fadd_rn(sx, c.sx); sy = __fadd_rn(sy, c.sy); sz = __fadd_rn(sz, c.sz);
        n += c.n; sr += c.sr; sg += c.sg; sb += c.sb;
    }
};

// keys of the cycle's points on the combined grid + whole-array digit histograms (for the sort plan)
__global__ void __launch_bounds__(kThreads) k_acc_key_pts(const float4* __restrict__ pts, const uint32_t* __restrict__ n_ptr,
                                                          float ix, float iy, float iz, uint64_t* __restrict__ keys,
                                                          uint32_t* __restrict__ vals, uint32_t* __restrict__ ghist) {
    __shared__ uint32_t sh[kMaxPasses * kRsBins];
    for (int i = threadIdx.x; i < kMaxPasses * kRsBins; i += kThreads) sh[i] = 0;
    __syncthreads();
    const uint32_t n = *n_ptr;
    const int lane = threadIdx.x & 31;
    for (uint32_t base = blockIdx.x * kThreads; base < n; base += gridDim.x * kThreads) {
        const uint32_t i = base + threadIdx.x;
        const bool valid = i < n;
        uint64_t k = 0;
        if (valid) {
            const float4 p = pts[i];
            k = abs_cell_key(p.x, p.y, __fadd_rn(p.z, 500.0f), ix, iy, iz);
            keys[i] = k;
            vals[i] = i;
        }
        const unsigned vm = __ballot_sync(kFull, valid);
        if (valid) {
#pragma unroll
            for (int p = 0; p < kMaxPasses; ++p) {
                const uint32_t d = (uint32_t)(k >> (8 * p)) & 255u;
                const unsigned peers = __match_any_sync(vm, d);
                if (lane == __ffs(peers) - 1) atomicAdd(&sh[p * kRsBins + d], (uint32_t)__popc(peers));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kMaxPasses * kRsBins; i += kThreads)
        if (sh[i]) atomicAdd(&ghist[i], sh[i]);
}

__global__ void __launch_bounds__(kThreads) k_acc_key_cells(const o3r_cell* __restrict__ cells, uint32_t n,
                                                            uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                            uint32_t* __restrict__ ghist) {
    __shared__ uint32_t sh[kMaxPasses * kRsBins];
    for (int i = threadIdx.x; i < kMaxPasses * kRsBins; i += kThreads) sh[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (uint32_t base = blockIdx.x * kThreads; base < n; base += gridDim.x * kThreads) {
        const uint32_t i = base + threadIdx.x;
        const bool valid = i < n;
        uint64_t k = 0;
        if (valid) { k = cells[i].key; keys[i] = k; vals[i] = i; }
        const unsigned vm = __ballot_sync(kFull, valid);
        if (valid) {
#pragma unroll
            for (int p = 0; p < kMaxPasses; ++p) {
                const uint32_t d = (uint32_t)(k >> (8 * p)) & 255u;
                const unsigned peers = __match_any_sync(vm, d);
                if (lane == __ffs(peers) - 1) atomicAdd(&sh[p * kRsBins + d], (uint32_t)__popc(peers));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kMaxPasses * kRsBins; i += kThreads)
        if (sh[i]) atomicAdd(&ghist[i], sh[i]);
}

struct AccArgs {
    const uint64_t* keys0; const uint64_t* keys1;
    const uint32_t* vals0; const uint32_t* vals1;
    const uint32_t* seg_off;      // [2] = {0, n}
    const SortPlan* plan;         // [1]
    uint32_t tiles_ub;
    // resident shard (sorted by key)
    const uint64_t* res_keys; float4* res_acc; uint4* res_rgb; uint32_t n_res;
};

__global__ void __launch_bounds__(kThreads) k_acc_heads(AccArgs A, uint32_t* __restrict__ head_cnt) {
    const uint32_t t = blockIdx.x;
    const uint32_t n = A.seg_off[1];
    uint32_t c = 0;
    if (t * kTileV < n) {
        const uint64_t* keys = A.plan[0].final_parity ? A.keys1 : A.keys0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t pos = t * kTileV + threadIdx.x * 4 + j;
            if (pos < n && (pos == 0 || keys[pos] != keys[pos - 1])) ++c;
        }
    }
    c = __reduce_add_sync(kFull, c);
    __shared__ uint32_t s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_c, c);
    __syncthreads();
    if (threadIdx.x == 0) head_cnt[t] = s_c;
}

// One thread per run: look the cell up in the resident shard, continue its sums in input order.
// Writes the cycle's cell list (sorted by key): ckey, cacc = {sx, sy, sz, n}, crgb = {sr, sg, sb, found_pos + 1}.
template <typename Items>
__global__ void __launch_bounds__(kThreads) k_acc_reduce(AccArgs A, Items items, const uint32_t* __restrict__ head_off,
                                                         uint64_t* __restrict__ ckey, float4* __restrict__ cacc,
                                                         uint4* __restrict__ crgb, uint32_t* __restrict__ n_new) {
    __shared__ uint32_t s_scan[34];
    const uint32_t t = blockIdx.x;
    const uint32_t n = A.seg_off[1];
    if (t * kTileV >= n) return;
    const int par = A.plan[0].final_parity;
    const uint64_t* keys = par ? A.keys1 : A.keys0;
    const uint32_t* vals = par ? A.vals1 : A.vals0;
    uint32_t flags = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t pos = t * kTileV + threadIdx.x * 4 + j;
        if (pos < n && (pos == 0 || keys[pos] != keys[pos - 1])) flags |= 1u << j;
    }
    uint32_t tot;
    uint32_t o = head_off[t] + block_excl_scan((uint32_t)__popc(flags), s_scan, tot);
    uint32_t fresh = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (!(flags & (1u << j))) continue;
        const uint32_t pos = t * kTileV + threadIdx.x * 4 + j;
        const uint64_t k = keys[pos];
        uint32_t lo = 0, hi = A.n_res;  // lower_bound in the resident keys
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (A.res_keys[mid] < k) lo = mid + 1; else hi = mid;
        }
        const bool found = lo < A.n_res && A.res_keys[lo] == k;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        uint32_t cn = 0, sr = 0, sg = 0, sb = 0;
        if (found) {
            const float4 a = A.res_acc[lo];
            const uint4 c = A.res_rgb[lo];
            sx = a.x; sy = a.y; sz = a.z; cn = __float_as_uint(a.w); sr = c.x; sg = c.y; sb = c.z;
        } else {
            ++fresh;
        }
        for (uint32_t q = pos; q < n && keys[q] == k; ++q) items.add(vals[q], sx, sy, sz, cn, sr, sg, sb);
        ckey[o] = k;
        cacc[o] = make_float4(sx, sy, sz, __uint_as_float(cn));
        crgb[o] = make_uint4(sr, sg, sb, found ? lo + 1 : 0u);
        ++o;
    }
    fresh = __reduce_add_sync(kFull, fresh);
    if ((threadIdx.x & 31) == 0 && fresh) atomicAdd(n_new, fresh);
}

// found cells: write the continued sums back in place; new cells: flag for compaction
__global__ void __launch_bounds__(kThreads) k_acc_update(uint32_t n_cyc, const float4* __restrict__ cacc,
                                                         const uint4* __restrict__ crgb, float4* __restrict__ res_acc,
                                                         uint4* __restrict__ res_rgb, uint32_t* __restrict__ new_cnt) {
    const uint32_t t = blockIdx.x;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t h = t * kTileV + threadIdx.x * 4 + j;
        if (h < n_cyc) {
            const uint4 r = crgb[h];
            if (r.w) {
                res_acc[r.w - 1] = cacc[h];
                res_rgb[r.w - 1] = make_uint4(r.x, r.y, r.z, 0u);
            } else {
                ++c;
            }
        }
    }
    c = __reduce_add_sync(kFull, c);
    __shared__ uint32_t s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_c, c);
    __syncthreads();
    if (threadIdx.x == 0) new_cnt[t] = s_c;
}

// new cells go to their merged position: q-th new cell -> q + lower_bound(res_keys, key)
__global__ void __launch_bounds__(kThreads) k_acc_place_new(uint32_t n_cyc, const uint64_t* __restrict__ ckey,
                                                            const float4* __restrict__ cacc, const uint4* __restrict__ crgb,
                                                            const uint32_t* __restrict__ new_off,
                                                            const uint64_t* __restrict__ res_keys, uint32_t n_res,
                                                            uint64_t* __restrict__ dst_keys, float4* __restrict__ dst_acc,
                                                            uint4* __restrict__ dst_rgb, uint64_t* __restrict__ new_keys) {
    __shared__ uint32_t s_scan[34];
    const uint32_t t = blockIdx.x;
    uint32_t flags = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t h = t * kTileV + threadIdx.x * 4 + j;
        if (h < n_cyc && crgb[h].w == 0) flags |= 1u << j;
    }
    uint32_t tot;
    uint32_t q = new_off[t] + block_excl_scan((uint32_t)__popc(flags), s_scan, tot);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (!(flags & (1u << j))) continue;
        const uint32_t h = t * kTileV + threadIdx.x * 4 + j;
        const uint64_t k = ckey[h];
        uint32_t lo = 0, hi = n_res;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (res_keys[mid] < k) lo = mid + 1; else hi = mid;
        }
        const uint32_t d = q + lo;
        dst_keys[d] = k;
        dst_acc[d] = cacc[h];
        const uint4 r = crgb[h];
        dst_rgb[d] = make_uint4(r.x, r.y, r.z, 0u);
        new_keys[q] = k;
        ++q;
    }
}

// resident cells shift right by the number of new cells with a smaller key
__global__ void __launch_bounds__(kThreads) k_acc_place_old(uint32_t n_res, const uint64_t* __restrict__ res_keys,
                                                            const float4* __restrict__ res_acc, const uint4* __restrict__ res_rgb,
                                                            const uint64_t* __restrict__ new_keys, uint32_t n_new,
                                                            uint64_t* __restrict__ dst_keys, float4* __restrict__ dst_acc,
                                                            uint4* __restrict__ dst_rgb) {
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n_res; i += gridDim.x * kThreads) {
        const uint64_t k = res_keys[i];
        uint32_t lo = 0, hi = n_new;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (new_keys[mid] < k) lo = mid + 1; else hi = mid;
        }
        const uint32_t d = i + lo;
        dst_keys[d] = k;
        dst_acc[d] = res_acc[i];
        dst_rgb[d] = res_rgb[i];
    }
}

// downsamplePtCloud(cloud_big, true) on the accumulators: count >= min_points -> centroid, z -= 500
__global__ void __launch_bounds__(kThreads) k_acc_emit_cnt(uint32_t n_res, const float4* __restrict__ res_acc,
                                                           uint32_t min_points, uint32_t* __restrict__ cnt) {
    const uint32_t t = blockIdx.x;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t i = t * kTileV + threadIdx.x * 4 + j;
        if (i < n_res && __float_as_uint(res_acc[i].w) >= min_points) ++c;
    }
    c = __reduce_add_sync(kFull, c);
    __shared__ uint32_t s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_c, c);
    __syncthreads();
    if (threadIdx.x == 0) cnt[t] = s_c;
}

__global__ void __launch_bounds__(kThreads) k_acc_emit(uint32_t n_res, const float4* __restrict__ res_acc,
                                                       const uint4* __restrict__ res_rgb, uint32_t min_points,
                                                       const uint32_t* __restrict__ off, float4* __restrict__ out) {
    __shared__ uint32_t s_scan[34];
    const uint32_t t = blockIdx.x;
    uint32_t flags = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t i = t * kTileV + threadIdx.x * 4 + j;
        if (i < n_res && __float_as_uint(res_acc[i].w) >= min_points) flags |= 1u << j;
    }
    uint32_t tot;
    uint32_t o = off[t] + block_excl_scan((uint32_t)__popc(flags), s_scan, tot);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (!(flags & (1u << j))) continue;
        const uint32_t i = t * kTileV + threadIdx.x * 4 + j;
        const float4 a = res_acc[i];
        const uint4 c = res_rgb[i];
        const float fn = (float)__float_as_uint(a.w);
        const uint32_t rgb = ((uint32_t)__fdiv_rn((float)c.x, fn) << 16) | ((uint32_t)__fdiv_rn((float)c.y, fn) << 8) |
                             (uint32_t)__fdiv_rn((float)c.z, fn);
        out[o++] = make_float4(__fdiv_rn(a.x, fn), __fdiv_rn(a.y, fn), __fsub_rn(__fdiv_rn(a.z, fn), 500.0f),
                               __uint_as_float(rgb));
    }
}

// ---- multi-GPU exchange helpers -------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t hash64(uint64_t x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

// owner of each cycle cell + per-owner histogram; keys32 = owner, vals = cell index (then one sort pass)
__global__ void __launch_bounds__(kThreads) k_owner(uint32_t n_cyc, const uint64_t* __restrict__ ckey, uint32_t world,
                                                    uint32_t* __restrict__ okeys, uint32_t* __restrict__ ovals,
                                                    uint32_t* __restrict__ counts) {
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n_cyc; i += gridDim.x * kThreads) {
        const uint32_t o = (uint32_t)(hash64(ckey[i]) % world);
        okeys[i] = o;
        ovals[i] = i;
        atomicAdd(&counts[o], 1u);
    }
}

__global__ void __launch_bounds__(kThreads) k_pack_cells(uint32_t n_cyc, const uint32_t* __restrict__ order,
                                                         const uint64_t* __restrict__ ckey, const float4* __restrict__ cacc,
                                                         const uint4* __restrict__ crgb, o3r_cell* __restrict__ out) {
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n_cyc; i += gridDim.x * kThreads) {
        const uint32_t h = order[i];
        const float4 a = cacc[h];
        const uint4 c = crgb[h];
        o3r_cell r;
        r.key = ckey[h];
        r.sx = a.x; r.sy = a.y; r.sz = a.z; r.n = __float_as_uint(a.w);
        r.sr = c.x; r.sg = c.y; r.sb = c.z; r.pad = 0;
        out[i] = r;
    }
}

// rigid transform of a resident point cloud in place (pose.cpp:353)
__global__ void __launch_bounds__(kThreads) k_transform_pts(float4* __restrict__ pts, size_t n, const float* __restrict__ Tm) {
    __shared__ float T[12];
    if (threadIdx.x < 12) T[threadIdx.x] = Tm[threadIdx.x];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        float4 p = pts[i];
        float x, y, z;
        xform(T, p.x, p.y, p.z, x, y, z);
        p.x = x; p.y = y; p.z = z;
        pts[i] = p;
    }
}

}  // namespace o3r
