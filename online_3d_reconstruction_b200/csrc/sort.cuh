// sort.cuh — hand-written segmented LSD radix sort of (key, u32 value) pairs, plus the exclusive-scan
// kernels the pipeline uses.  No Thrust / CUB.
//
// 8-bit digits; one pass = k_rs_hist (per-tile digit histogram) -> k_rs_scan (digit-major exclusive
// scan per segment) -> k_rs_scatter (stable in-tile ranking with warp match_any, exchange through
// shared memory so that every digit run leaves the CTA as one coalesced burst).  Stability is what
// makes the later per-voxel sums run in input order (bit-exact with the oracle's stable order).
//
// Segments (one per frame for the per-frame grid, a single one for the combined grid) are described
// by a device array seg_off[S+1]; launches are sized by a host upper bound and surplus CTAs exit.
// A per-segment SortPlan lets passes whose digit is constant over the segment be skipped entirely
// (the 64-bit absolute cell keys of the combined grid have 4-6 such digits).
#pragma once
#include "common.cuh"

namespace o3r {

constexpr int kRsItems = 16;
constexpr int kRsTile = kThreads * kRsItems;  // 4096 pairs per CTA
constexpr int kRsBins = 256;
constexpr int kMaxPasses = 8;

struct SortPlan {
    uint8_t active[kMaxPasses];
    uint8_t in_parity[kMaxPasses];   // which ping-pong buffer pass p reads
    uint32_t final_parity;           // buffer holding the sorted result
    uint32_t n_active;
};

// ---- generic single-segment exclusive scan (one CTA, coalesced, carry across iterations) ----------------
__global__ void __launch_bounds__(kThreads) k_scan_u32(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                       uint32_t n, uint32_t* __restrict__ total_out) {
    __shared__ uint32_t sm[34];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += kThreads * 4) {
        const uint32_t i = base + threadIdx.x * 4;
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (i + j < n) ? in[i + j] : 0u;
        uint32_t tot;
        uint32_t off = block_excl_scan(v[0] + v[1] + v[2] + v[3], sm, tot) + carry;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i + j < n) out[i + j] = off;
            off += v[j];
        }
        carry += tot;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// ---- per-segment plan from whole-segment digit histograms ghist[S][passes][256] ---------------------------
__global__ void k_rs_plan(const uint32_t* __restrict__ ghist, const uint32_t* __restrict__ seg_off, int passes,
                          SortPlan* __restrict__ plan) {
    __shared__ int s_trivial[kMaxPasses];
    const int s = blockIdx.x;
    const uint32_t n = seg_off[s + 1] - seg_off[s];
    if (threadIdx.x < kMaxPasses) s_trivial[threadIdx.x] = 0;
    __syncthreads();
    for (int p = 0; p < passes; ++p)
        if (ghist[((size_t)s * passes + p) * kRsBins + threadIdx.x] == n) s_trivial[p] = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        SortPlan pl;
        uint32_t par = 0, na = 0;
        for (int p = 0; p < kMaxPasses; ++p) {
            const bool act = p < passes && n > 0 && !s_trivial[p];
            pl.active[p] = act;
            pl.in_parity[p] = (uint8_t)par;
            if (act) { par ^= 1u; ++na; }
        }
        pl.final_parity = par;
        pl.n_active = na;
        plan[s] = pl;
    }
}

template <typename KeyT>
__device__ __forceinline__ uint32_t rs_digit(KeyT k, int shift) { return (uint32_t)(k >> shift) & 255u; }

// ---- pass step 1: per-tile histogram -------------------------------------------------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) k_rs_hist(const KeyT* __restrict__ keys0, const KeyT* __restrict__ keys1,
                                                      const uint32_t* __restrict__ seg_off,
                                                      const SortPlan* __restrict__ plan, int pass, uint32_t tiles_ub,
                                                      uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[kRsBins];
    const int s = blockIdx.y;
    const uint32_t t = blockIdx.x;
    if (!plan[s].active[pass]) return;
    const uint32_t beg = seg_off[s], n = seg_off[s + 1] - beg;
    const uint32_t nt = (n + kRsTile - 1) / kRsTile;
    if (t >= nt) return;
    const KeyT* keys = (plan[s].in_parity[pass] ? keys1 : keys0) + beg;
    const int shift = pass * 8, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    sh[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kRsItems; ++r) {
        const uint32_t i = t * kRsTile + warp * (32 * kRsItems) + r * 32 + lane;
        const bool valid = i < n;
        const unsigned vm = __ballot_sync(kFull, valid);
        if (valid) {
            const uint32_t d = rs_digit(keys[i], shift);
            const unsigned peers = __match_any_sync(vm, d);
            if (lane == __ffs(peers) - 1) atomicAdd(&sh[d], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    hist[(size_t)s * kRsBins * tiles_ub + (size_t)threadIdx.x * nt + t] = sh[threadIdx.x];
}

// ---- pass step 2: exclusive scan of the segment's [256][nt] histogram, digit-major ----------------------------
__global__ void __launch_bounds__(kThreads) k_rs_scan(const uint32_t* __restrict__ seg_off,
                                                      const SortPlan* __restrict__ plan, int pass, uint32_t tiles_ub,
                                                      uint32_t* __restrict__ hist) {
    __shared__ uint32_t sm[34];
    const int s = blockIdx.x;
    if (!plan[s].active[pass]) return;
    const uint32_t n = seg_off[s + 1] - seg_off[s];
    const uint32_t nt = (n + kRsTile - 1) / kRsTile;
    const uint32_t len = nt * kRsBins;
    uint32_t* h = hist + (size_t)s * kRsBins * tiles_ub;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < len; base += kThreads * 4) {
        const uint32_t i = base + threadIdx.x * 4;
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (i + j < len) ? h[i + j] : 0u;
        uint32_t tot;
        uint32_t off = block_excl_scan(v[0] + v[1] + v[2] + v[3], sm, tot) + carry;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i + j < len) h[i + j] = off;
            off += v[j];
        }
        carry += tot;
    }
}

// ---- pass step 3: stable rank + scatter -------------------------------------------------------------------------
// Dynamic shared memory: cnt[kWarps][256] | dstart[256] | gbase[256] | keys[4096] | vals[4096]
template <typename KeyT>
constexpr size_t rs_scatter_smem() {
    return (size_t)(kWarps * kRsBins + 2 * kRsBins + 34) * 4 + (size_t)kRsTile * (sizeof(KeyT) + 4);
}

template <typename KeyT>
__global__ void __launch_bounds__(kThreads) k_rs_scatter(KeyT* __restrict__ keys0, KeyT* __restrict__ keys1,
                                                         uint32_t* __restrict__ vals0, uint32_t* __restrict__ vals1,
                                                         const uint32_t* __restrict__ seg_off,
                                                         const SortPlan* __restrict__ plan, int pass, uint32_t tiles_ub,
                                                         const uint32_t* __restrict__ hist, int iota_first) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    const int s = blockIdx.y;
    const uint32_t t = blockIdx.x;
    const SortPlan pl = plan[s];
    if (!pl.active[pass]) return;
    const uint32_t beg = seg_off[s], n = seg_off[s + 1] - beg;
    const uint32_t nt = (n + kRsTile - 1) / kRsTile;
    if (t >= nt) return;
    KeyT* s_keys = reinterpret_cast<KeyT*>(rs_smem);
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(rs_smem + (size_t)kRsTile * sizeof(KeyT));
    uint32_t* cnt = s_vals + kRsTile;            // [kWarps][256]
    uint32_t* dstart = cnt + kWarps * kRsBins;   // [256] tile-local start of each digit run
    uint32_t* gbase = dstart + kRsBins;          // [256] segment-relative destination of each digit run
    uint32_t* s_scan = gbase + kRsBins;          // [34]

    const int par = pl.in_parity[pass];
    const KeyT* kin = (par ? keys1 : keys0) + beg;
    const uint32_t* vin = (par ? vals1 : vals0) + beg;
    KeyT* kout = (par ? keys0 : keys1) + beg;
    uint32_t* vout = (par ? vals0 : vals1) + beg;
    const bool iota = iota_first && pl.in_parity[pass] == 0 && pass == 0;  // values are the global index
    (void)iota;
    const int shift = pass * 8, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;

#pragma unroll
    for (int w = 0; w < kWarps; ++w) cnt[w * kRsBins + threadIdx.x] = 0;
    gbase[threadIdx.x] = hist[(size_t)s * kRsBins * tiles_ub + (size_t)threadIdx.x * nt + t];
    __syncthreads();

    KeyT key[kRsItems];
    uint32_t val[kRsItems];
    uint16_t rank[kRsItems];
    const uint32_t wbase = t * kRsTile + warp * (32 * kRsItems);
    const uint32_t ntile = min((uint32_t)kRsTile, n - t * kRsTile);
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? kin[i] : ~(KeyT)0;
        val[r] = valid ? (iota_first && pass == 0 ? beg + i : vin[i]) : 0u;
    }
    uint32_t* wc = cnt + warp * kRsBins;
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        // padding items carry digit 255 and sit at the very end of the tile order, so they never
        // precede a real item inside any digit run
        const uint32_t d = rs_digit(key[r], shift);
        const unsigned peers = __match_any_sync(kFull, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader) { old = wc[d]; wc[d] = old + __popc(peers); }
        old = __shfl_sync(kFull, old, leader);
        rank[r] = (uint16_t)(old + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over warps, tile total
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const uint32_t c = cnt[w * kRsBins + threadIdx.x];
        cnt[w * kRsBins + threadIdx.x] = run;
        run += c;
    }
    uint32_t tot;
    const uint32_t ds = block_excl_scan(run, s_scan, tot);
    dstart[threadIdx.x] = ds;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const uint32_t d = rs_digit(key[r], shift);
        const uint32_t pos = dstart[d] + wc[d] + rank[r];
        s_keys[pos] = key[r];
        s_vals[pos] = val[r];
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < ntile; i += kThreads) {
        const KeyT k = s_keys[i];
        const uint32_t d = rs_digit(k, shift);
        const uint32_t g = gbase[d] + (i - dstart[d]);
        kout[g] = k;
        vout[g] = s_vals[i];
    }
}

}  // namespace o3r
