// sort.cuh — hand-written segmented LSD radix sort of (key, u32 value) pairs, plus the exclusive-scan
// kernels the pipeline uses.  No Thrust / CUB.
//
// 8-bit digits.  k_rs_ghist reads the keys once and builds every pass's whole-segment digit histogram;
// each pass is then ONE kernel (k_rs_onesweep): stable in-tile ranking with warp match_any, tile offsets
// by decoupled look-back, exchange through shared memory so that every digit run leaves the CTA as one
// coalesced burst.  Stability is what makes the later per-voxel sums run in input order (bit-exact with
// the oracle's stable order).
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
                          SortPlan* __restrict__ plan, const GridParams* __restrict__ grids) {
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
            const bool act = p < passes && n > 0 && !s_trivial[p] && !(grids && grids[s].passthrough);
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

// shared-memory histogram update by the lanes of `vm`: one aggregated atomic when the whole warp agrees on the
// digit (the constant high digits that would otherwise serialise), plain atomics otherwise
__device__ __forceinline__ void hist_add(uint32_t* sh, uint32_t d, unsigned vm, int lane) {
    int uniform;
    __match_all_sync(vm, d, &uniform);
    if (uniform) {
        if (lane == __ffs(vm) - 1) atomicAdd(&sh[d], (uint32_t)__popc(vm));
    } else {
        atomicAdd(&sh[d], 1u);
    }
}

// ---- whole-segment digit histograms for every pass in one read of the keys ------------------------------------
// ghist[S][PASSES][256]; must be zeroed by the caller.
template <typename KeyT, int PASSES>
__global__ void __launch_bounds__(kThreads) k_rs_ghist(const KeyT* __restrict__ keys, const uint32_t* __restrict__ seg_off,
                                                       uint32_t* __restrict__ ghist) {
    __shared__ uint32_t sh[PASSES * kRsBins];
    const int s = blockIdx.y;
    const uint32_t t = blockIdx.x;
    const uint32_t beg = seg_off[s], n = seg_off[s + 1] - beg;
    if (t * kRsTile >= n) return;
    for (int i = threadIdx.x; i < PASSES * kRsBins; i += kThreads) sh[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 2
    for (int r = 0; r < kRsItems; ++r) {
        const uint32_t i = t * kRsTile + warp * (32 * kRsItems) + r * 32 + lane;
        const bool valid = i < n;
        const unsigned vm = __ballot_sync(kFull, valid);
        if (valid) {
            const KeyT k = keys[beg + i];
#pragma unroll
            for (int p = 0; p < PASSES; ++p) hist_add(sh + p * kRsBins, rs_digit(k, 8 * p), vm, lane);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PASSES * kRsBins; i += kThreads)
        if (sh[i]) atomicAdd(&ghist[(size_t)s * PASSES * kRsBins + i], sh[i]);
}

// ---- one radix pass in one kernel: stable in-tile rank, decoupled look-back for the tile's digit offsets ---------
// ("onesweep"): every tile publishes its per-digit counts in status[seg][tile][256] (2 flag bits + 30 count
// bits in one word), then each of the 256 threads walks back over the predecessors' words of ITS digit until it
// meets an inclusive prefix.  Tiles take their index from an atomic ticket so a predecessor is always resident
// or finished before anyone waits on it.
// Dynamic shared memory: keys[4096] | vals[4096] | cnt[kWarps][256] | dstart[256] | gbase[256] | scan[34]
template <typename KeyT>
constexpr size_t rs_scatter_smem() {
    return (size_t)(kWarps * kRsBins + 2 * kRsBins + 34 + 2) * 4 + (size_t)kRsTile * (sizeof(KeyT) + 4);
}

constexpr uint32_t kStLocal = 1u << 30, kStGlobal = 2u << 30, kStMask = (1u << 30) - 1u;

// flag + payload live in ONE word, so relaxed gpu-scope accesses are enough (no fence, no system scope)
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename KeyT>
__global__ void __launch_bounds__(kThreads, sizeof(KeyT) == 4 ? 4 : 3) k_rs_onesweep(KeyT* __restrict__ keys0, KeyT* __restrict__ keys1,
                                                          uint32_t* __restrict__ vals0, uint32_t* __restrict__ vals1,
                                                          const uint32_t* __restrict__ seg_off,
                                                          const SortPlan* __restrict__ plan, int pass, int passes,
                                                          uint32_t tiles_ub, const uint32_t* __restrict__ ghist,
                                                          uint32_t* __restrict__ status, uint32_t* __restrict__ ticket,
                                                          int iota_first, const float4* __restrict__ gsrc,
                                                          float4* __restrict__ gdst) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    KeyT* s_keys = reinterpret_cast<KeyT*>(rs_smem);
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(rs_smem + (size_t)kRsTile * sizeof(KeyT));
    uint32_t* cnt = s_vals + kRsTile;            // [kWarps][256]
    uint32_t* dstart = cnt + kWarps * kRsBins;   // [256] tile-local start of each digit run
    uint32_t* gbase = dstart + kRsBins;          // [256] segment-relative destination of each digit run
    uint32_t* s_scan = gbase + kRsBins;          // [34]
    uint32_t* s_ticket = s_scan + 34;

    if (threadIdx.x == 0) *s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t lin = *s_ticket;
    const int s = lin / tiles_ub;
    const uint32_t t = lin - (uint32_t)s * tiles_ub;
    const SortPlan pl = plan[s];
    if (!pl.active[pass]) return;
    const uint32_t beg = seg_off[s], n = seg_off[s + 1] - beg;
    const uint32_t nt = (n + kRsTile - 1) / kRsTile;
    if (t >= nt) return;

    const int par = pl.in_parity[pass];
    const KeyT* kin = (par ? keys1 : keys0) + beg;
    const uint32_t* vin = (par ? vals1 : vals0) + beg;
    KeyT* kout = (par ? keys0 : keys1) + beg;
    uint32_t* vout = (par ? vals0 : vals1) + beg;
    bool first = true;
    for (int q = 0; q < pass; ++q) first = first && !pl.active[q];
    bool last = true;
    for (int q = pass + 1; q < passes; ++q) last = last && !pl.active[q];
    // in the last active pass the 16-byte records the values point at can be delivered in sorted order
    // (gdst[beg + rank] = gsrc[value]) instead of the values: 16 independent gathers per thread
    const bool gather = last && gdst != nullptr;
    const bool iota = iota_first && first;   // values are the global element index, not read from memory
    const int shift = pass * 8, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;

#pragma unroll
    for (int w = 0; w < kWarps; ++w) cnt[w * kRsBins + threadIdx.x] = 0;
    // exclusive scan of the segment's digit histogram = where each digit's run starts in the segment
    uint32_t tot;
    const uint32_t dbase = block_excl_scan(ghist[((size_t)s * passes + pass) * kRsBins + threadIdx.x], s_scan, tot);

    KeyT key[kRsItems];
    uint32_t rk[kRsItems / 2];   // two 16-bit in-warp ranks per word
    const uint32_t wbase = t * kRsTile + warp * (32 * kRsItems);
    const uint32_t ntile = min((uint32_t)kRsTile, n - t * kRsTile);
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        key[r] = (i < n) ? kin[i] : ~(KeyT)0;
    }
    uint32_t* wc = cnt + warp * kRsBins;
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        // padding items carry digit 255 and sit at the very end of the tile order, so they never
        // precede a real item inside any digit run
        const uint32_t d = rs_digit(key[r], shift);
        const unsigned peers = __match_any_sync(kFull, d);
        const uint32_t old = wc[d];                   // every peer reads the counter (broadcast) ...
        __syncwarp();
        if ((peers & lt) == 0u) wc[d] = old + __popc(peers);   // ... then the lowest peer bumps it
        __syncwarp();
        const uint32_t rnk = old + __popc(peers & lt);
        if (r & 1) rk[r >> 1] |= rnk << 16; else rk[r >> 1] = rnk;
    }
    // values are only needed for the exchange below: load them now so the latency hides behind the scans
    uint32_t val[kRsItems];
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        val[r] = (i < n) ? (iota ? beg + i : vin[i]) : 0u;
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
    // real items of this digit in this tile (padding only ever inflates digit 255)
    uint32_t real = run;
    if (threadIdx.x == kRsBins - 1) real -= (kRsTile - ntile);
    // publish, look back
    uint32_t* st = status + ((size_t)s * tiles_ub + t) * kRsBins + threadIdx.x;
    uint32_t prefix = 0;
    if (t == 0) {
        st_volatile_u32(st, kStGlobal | real);
    } else {
        st_volatile_u32(st, kStLocal | real);
        const uint32_t* pst = st - kRsBins;
        for (uint32_t back = t; back > 0; --back, pst -= kRsBins) {
            uint32_t v;
            while (((v = ld_volatile_u32(pst)) >> 30) == 0u) __nanosleep(32);
            prefix += v & kStMask;
            if ((v >> 30) == 2u) break;
        }
        st_volatile_u32(st, kStGlobal | (prefix + real));
    }
    const uint32_t ds = block_excl_scan(run, s_scan, tot);
    dstart[threadIdx.x] = ds;
    gbase[threadIdx.x] = dbase + prefix - ds;   // destination of tile-local position i of this digit: gbase[d] + i
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const uint32_t d = rs_digit(key[r], shift);
        const uint32_t pos = dstart[d] + wc[d] + ((rk[r >> 1] >> ((r & 1) * 16)) & 0xffffu);
        s_keys[pos] = key[r];
        s_vals[pos] = val[r];
    }
    __syncthreads();
    if (gather) {
        float4* go = gdst + beg;
#pragma unroll 4
        for (uint32_t i = threadIdx.x; i < ntile; i += kThreads) {
            const KeyT k = s_keys[i];
            const uint32_t g = gbase[rs_digit(k, shift)] + i;
            kout[g] = k;
            go[g] = gsrc[s_vals[i]];
        }
    } else {
        for (uint32_t i = threadIdx.x; i < ntile; i += kThreads) {
            const KeyT k = s_keys[i];
            const uint32_t g = gbase[rs_digit(k, shift)] + i;
            kout[g] = k;
            vout[g] = s_vals[i];
        }
    }
}

}  // namespace o3r
