// o3r_api.cu — C ABI (include/o3r.h) over the sm_100a kernels.  No CPU fallback: every compute entry
// point launches CUDA kernels on the context's stream and fails with O3R_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include <chrono>

#include "blur.cuh"
#include "common.cuh"
#include "prepass.cuh"
#include "prereduce.cuh"
#include "sor.cuh"
#include "sort.cuh"
#include "stage_a.cuh"
#include "voxel.cuh"

using namespace o3r;

namespace {

std::string g_create_err;

inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct Tr {
    bool on; double t; const char* where;
    explicit Tr(const char* w) : on(getenv("O3R_TRACE") != nullptr), t(now_ms()), where(w) {}
    void mark(const char* what) { if (!on) return; const double n = now_ms(); fprintf(stderr, "[o3r trace] %s/%s %.3f ms\n", where, what, n - t); t = n; }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
    cudaError_t ensure(size_t bytes, cudaStream_t st = nullptr, size_t preserve = 0) {
        if (bytes <= cap) return cudaSuccess;
        size_t want = std::max(bytes, cap + cap / 2);
        want = (want + 255) & ~(size_t)255;
        void* np = nullptr;
        cudaError_t e = cudaMalloc(&np, want);
        if (e != cudaSuccess) return e;
        if (p && preserve) {
            e = cudaMemcpyAsync(np, p, preserve, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) return e;
            cudaStreamSynchronize(st);
        }
        if (p) cudaFree(p);
        p = np;
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

enum { CNT_PTS = 0, CNT_VOX = 1, CNT_CYC = 2, CNT_NEW = 3, CNT_EMIT = 4, CNT_NEWSCAN = 5, CNT_BASE = 6, CNT_NRES = 7, CNT_ZERO = 8, CNT_PART = 9, CNT_PARTCHUNK = 10,
       CNT_CELLBB = 16, CNT_N = 32 };

}  // namespace

struct o3r_ctx {
    o3r_params p;
    std::mutex mu;
    std::string err;
    uint64_t launches = 0;
    cudaStream_t st = nullptr, st_copy = nullptr, st_copy2 = nullptr;   // two copy streams: the per-copy set-up gaps of one hide behind the other
    cudaEvent_t ev_copy2 = nullptr;
    std::vector<cudaEvent_t> chunk_ev;
    int chunk_frames = 10, chunk_frames_dev = 1 << 30;
    cudaEvent_t chunk_event(size_t i) {
        while (chunk_ev.size() <= i) {
            cudaEvent_t e;
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            chunk_ev.push_back(e);
        }
        return chunk_ev[i];
    }
    int nx = 0, ny = 0;
    uint32_t npix = 0;
    int canon = 0;
    float leaf_f = 0, inv_f = 0, leaf_c = 0, inv_c = 0, inv_cz = 0;
    int defer_merge = 0;
    DevBuf lut_r, lut_z;
    // input staging (host-pointer entry points)
    // double-buffered: a prefetch of the next cycle's inputs fills one set while the kernels read the other
    struct Staging { DevBuf disp, bgr, labels, coef, kp; } stg[2];
    struct Prefetch {                   // one record per staging set
        bool valid = false;
        int n = 0, disp_type = 0;
        uint64_t seq = 0;               // issue order (the older record is recycled when both are pending)
        std::vector<const void*> sig;   // host pointers of the prefetched frames
        cudaEvent_t ev = nullptr;
    } prefetch[2];
    uint64_t prefetch_seq = 0;
    // A prefetch request is only RECORDED by o3r_frames_prefetch; its ~2 memcpy calls per frame are issued by the next
    // frame-path call right after that call's kernels are queued (the GPU computes while the host issues copies)
    // instead of in front of them with the GPU idle.
    struct Deferred { bool pending = false; std::vector<o3r_frame> frames; int disp_type = 0; } deferred;
    // a staging set that holds no pending prefetch (inputs of finished calls are free: every frame-path call returns
    // only after its last input-reading kernel has completed)
    int busy_set = -1;   // staging set the queued kernels of the running frame-path call still read
    int free_stage_set() {
        if (busy_set >= 0) { prefetch[busy_set ^ 1].valid = false; return busy_set ^ 1; }
        if (!prefetch[0].valid) return 0;
        if (!prefetch[1].valid) return 1;
        const int s = prefetch[0].seq < prefetch[1].seq ? 0 : 1;
        prefetch[s].valid = false;
        return s;
    }
    DevBuf d_disp;   // scratch plane of o3r_blur_u8
    DevBuf d_frames, d_blur, d_blurjobs;
    DevBuf pp_labels, pp_disp, pp_sums, pp_rows, pp_coef;   // pre-pass scratch (plane fit / variance gate)
    DevBuf bil_lut;            // bilateral LUTs (colour weights, space weights, per-row tap extents) of bil_kernel
    int bil_kernel = -1, bil_radius = 0, bil_maxk = 0;
    // per-batch work buffers
    DevBuf tile_cnt, tile_off, bbox, frame_off, grids, counters, pts, sortbuf, hist, plan_all, plan_v2, ghist;
    DevBuf head_cnt, head_off, vox, vox_off, seg2, tmat, mask, runwork, spts;
    uint32_t* h_counters = nullptr;   // pinned
    uint32_t* h_offs = nullptr;       // pinned, frame offsets readback
    size_t h_offs_cap = 0;
    // last batch
    int last_n = 0;
    size_t last_total = 0;
    bool last_is_vox = false;
    bool last_has_cellbb = false;
    int last_cellbb[6] = {0, 0, 0, 0, 0, 0};
    std::vector<uint32_t> last_off;
    // resident cloud: accumulators (ACCUMULATE) ...
    DevBuf res_keys[2], res_acc[2], res_rgb[2];
    int res_cur = 0;
    // The exact resident cell count lives on the device (counters[CNT_NRES]); the host keeps an upper bound that is
    // enough to size buffers and launches, and tightens it from an asynchronous read-back of the previous merge.
    size_t n_res_ub = 0;
    bool n_res_exact = true;
    uint32_t* h_nres = nullptr;       // pinned
    cudaEvent_t ev_nres = nullptr;
    size_t n_cyc_ub = 0;
    DevBuf ckey, cacc, crgb, new_cnt, new_off, new_keys, okeys;
    DevBuf sor_hard, sor_pts, sor_off, sor_dist, sor_grids, sor_pgrids, sor_rows, sor_thr, sor_skeys, sor_svals, sor_cnt, sor_cntoff;   // SOR scratch
    DevBuf partials, pr_status;   // TILED mode: the batch's tile partials (o3r_cell) and the look-back words
    size_t last_partials = 0;
    bool last_has_partials = false;
    uint32_t n_cyc = 0;
    // ... or points (RETAIN / dont_downsample)
    DevBuf cloud;
    size_t n_cloud = 0;

    // optional per-kernel event timing (bench.py's roofline leg)
    struct ProfRec { const char* name; cudaEvent_t a, b; };
    bool profiling = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t ev_get() {
        if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }

    bool retain() const { return p.dont_downsample || p.merge_mode == O3R_MERGE_RETAIN; }
    // tile pre-reduction is skipped while it does not reduce (probed again every 16th batch)
    unsigned tiled_poor = 0, tiled_batches = 0;
    bool tiled_now = false;
    bool tiled() const { return !p.dont_downsample && p.merge_mode == O3R_MERGE_ACCUMULATE_TILED; }
    int fail(int code, const std::string& m) { err = m; return code; }
    int fail_cuda(cudaError_t e, const char* what, int line) {
        char buf[512];
        snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s [o3r_api.cu:%d]", (int)e, cudaGetErrorString(e), what, line);
        err = buf;
        return O3R_ERR_CUDA;
    }
};

#define CU(call)                                                              \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) return ctx->fail_cuda(e_, #call, __LINE__);    \
    } while (0)

#define LAUNCH_N(name, kernel, grid, block, smem, ...)                        \
    do {                                                                      \
        cudaEvent_t pa_ = nullptr, pb_ = nullptr;                             \
        if (ctx->profiling) {                                                 \
            pa_ = ctx->ev_get(); pb_ = ctx->ev_get();                         \
            cudaEventRecord(pa_, ctx->st);                                    \
        }                                                                     \
        kernel<<<grid, block, smem, ctx->st>>>(__VA_ARGS__);                  \
        ++ctx->launches;                                                      \
        cudaError_t e_ = cudaGetLastError();                                  \
        if (e_ != cudaSuccess) return ctx->fail_cuda(e_, #kernel, __LINE__);  \
        if (pa_) {                                                            \
            cudaEventRecord(pb_, ctx->st);                                    \
            ctx->prof.push_back({name, pa_, pb_});                            \
        }                                                                     \
    } while (0)
#define LAUNCH(kernel, grid, block, smem, ...) LAUNCH_N(#kernel, kernel, grid, block, smem, __VA_ARGS__)

namespace {

inline uint32_t cdiv(size_t a, size_t b) { return (uint32_t)((a + b - 1) / b); }

// Small host->device payloads (frame descriptors, segment bounds, plans) travel as KERNEL PARAMETERS, not through the
// copy engine: a pageable cudaMemcpyAsync on the compute stream would queue behind a 184 MB input prefetch that
// occupies the H2D engine, and the whole cycle would wait for it.
struct SmallBlob { uint32_t w[960]; };   // 3840 bytes (kernel parameters are limited to 4 KB)
__global__ void k_put_blob(uint32_t* __restrict__ dst, SmallBlob b, int n_words) {
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = b.w[i];
}
int upload_small(o3r_ctx* ctx, void* dst, const void* src, size_t bytes) {
    const size_t cap = sizeof(SmallBlob);
    for (size_t at = 0; at < bytes; at += cap) {
        const size_t nb = std::min(cap, bytes - at);
        SmallBlob b;
        memcpy(b.w, (const char*)src + at, nb);
        if (nb % 4) memset((char*)b.w + nb, 0, 4 - nb % 4);
        LAUNCH(k_put_blob, 1, 256, 0, reinterpret_cast<uint32_t*>((char*)dst + at), b, (int)((nb + 3) / 4));
    }
    return O3R_OK;
}
// Zero fills run as kernels on the compute stream for the same reason: cudaMemsetAsync may be served by a copy engine
// and then queues behind an input prefetch.  `bytes` and `dst` are multiples of 4 (all callers clear u32 tables).
__global__ void __launch_bounds__(kThreads) k_zero(uint32_t* __restrict__ dst, size_t n_words) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    const size_t n4 = ((uintptr_t)dst % 16 == 0) ? n_words / 4 : 0;
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (size_t j = i; j < n4; j += stride) d4[j] = make_uint4(0u, 0u, 0u, 0u);
    for (size_t j = n4 * 4 + i; j < n_words; j += stride) dst[j] = 0u;
}
int zero_fill(o3r_ctx* ctx, void* dst, size_t bytes) {
    if (bytes == 0) return O3R_OK;
    const size_t words = (bytes + 3) / 4;
    const uint32_t g = (uint32_t)std::min<size_t>((words / 4 + kThreads - 1) / kThreads + 1, 148 * 8);
    LAUNCH(k_zero, g, kThreads, 0, reinterpret_cast<uint32_t*>(dst), words);
    return O3R_OK;
}
#define ZERO(ptr, bytes) do { int rcz_ = zero_fill(ctx, (ptr), (bytes)); if (rcz_) return rcz_; } while (0)
inline size_t disp_elem(int t) { return t == O3R_DISP_U8 ? 1 : t == O3R_DISP_U16 ? 2 : t == O3R_DISP_F32 ? 4 : 8; }

int read_counters(o3r_ctx* ctx) {
    CU(cudaMemcpyAsync(ctx->h_counters, ctx->counters.p, CNT_N * 4, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return O3R_OK;
}

// LUTs of cv::bilateralFilter(src, dst, d = k, sigmaColor = 2k, sigmaSpace = k/2) (pose_functions.cpp:1044), computed on the
// host with std::exp exactly as OpenCV 3.1 bilateralFilter_8u (and the oracle) do, cached per kernel size.
int bilateral_lut(o3r_ctx* ctx, int k, BilateralLut* out) {
    if (ctx->bil_kernel != k) {
        double sigma_color = (double)(k * 2), sigma_space = (double)(k / 2);
        if (sigma_color <= 0) sigma_color = 1;
        if (sigma_space <= 0) sigma_space = 1;
        const double gc = -0.5 / (sigma_color * sigma_color), gs = -0.5 / (sigma_space * sigma_space);
        const int radius = std::max(k <= 0 ? (int)std::lrint(sigma_space * 1.5) : k / 2, 1);
        std::vector<float> buf(256);
        for (int i = 0; i < 256; ++i) buf[i] = (float)std::exp(i * i * gc);
        std::vector<int> jm(2 * radius + 1, -1);
        for (int i = -radius; i <= radius; ++i)
            for (int j = -radius; j <= radius; ++j) {
                const double r = std::sqrt((double)i * i + (double)j * j);
                if (r > radius) continue;
                buf.push_back((float)std::exp(r * r * gs));
                jm[i + radius] = std::max(jm[i + radius], j);   // the mask is symmetric in j
            }
        const int maxk = (int)buf.size() - 256;
        const size_t bytes = buf.size() * 4 + jm.size() * 4;
        std::vector<unsigned char> blob(bytes);
        memcpy(blob.data(), buf.data(), buf.size() * 4);
        memcpy(blob.data() + buf.size() * 4, jm.data(), jm.size() * 4);
        CU(ctx->bil_lut.ensure(bytes));
        { int rcu = upload_small(ctx, ctx->bil_lut.p, blob.data(), bytes); if (rcu) return rcu; }
        ctx->bil_kernel = k; ctx->bil_radius = radius; ctx->bil_maxk = maxk;
        CU(cudaFuncSetAttribute(k_bilateral, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bilateral_smem(radius, maxk)));
    }
    out->color_w = ctx->bil_lut.as<float>();
    out->space_w = out->color_w + 256;
    out->jmax = reinterpret_cast<const int*>(out->space_w + ctx->bil_maxk);
    out->radius = ctx->bil_radius; out->maxk = ctx->bil_maxk;
    return O3R_OK;
}

// ---- radix sort driver -----------------------------------------------------------------------------------------
// The caller provides ghist [n_seg][passes][256] (already filled) and plan.
template <typename KeyT>
int sort_pairs(o3r_ctx* ctx, KeyT* k0, KeyT* k1, uint32_t* v0, uint32_t* v1, const uint32_t* seg_off, int n_seg,
               size_t per_seg_cap, const SortPlan* plan, int passes, int iota_first, const uint32_t* ghist,
               int ghist_is_prefix = 1) {
    const uint32_t tiles_ub = std::max(1u, cdiv(per_seg_cap, kRsTile));
    const size_t st_words = (size_t)n_seg * tiles_ub * kRsBins;
    // status words for every pass + one ticket per pass, cleared with one memset
    CU(ctx->hist.ensure((st_words * passes + 64) * 4));
    ZERO(ctx->hist.p, (st_words * passes + 64) * 4);
    uint32_t* status = ctx->hist.as<uint32_t>();
    uint32_t* tickets = status + st_words * passes;
    const uint32_t grid = tiles_ub * (uint32_t)n_seg;
    for (int p = 0; p < passes; ++p)
        LAUNCH_N(sizeof(KeyT) == 4 ? "k_rs_onesweep_u32" : "k_rs_onesweep_u64", (k_rs_onesweep<KeyT>), grid, kThreads,
                 rs_scatter_smem<KeyT>(), k0, k1, v0, v1, seg_off, plan, p, tiles_ub, ghist,
                 status + st_words * p, tickets + p, iota_first, ghist_is_prefix);
    return O3R_OK;
}

// ---- engine 1: VoxelGrid on segments whose leaf indices already sit in sortbuf keys0 ----------------------------
struct SortU32 { uint32_t *k0, *k1, *v0, *v1; };

int carve_sort_u32(o3r_ctx* ctx, size_t n, SortU32& s) {
    const size_t n4 = (n + 63) & ~(size_t)63;
    CU(ctx->sortbuf.ensure(n4 * 16));
    s.k0 = ctx->sortbuf.as<uint32_t>();
    s.k1 = s.k0 + n4; s.v0 = s.k1 + n4; s.v1 = s.v0 + n4;
    return O3R_OK;
}

// digit layout, whole-segment histograms and pass plan of a segmented u32 sort whose keys sit in sb.k0
// (grids[s].key_bits = live key bits of segment s); leaves the plan in ctx->plan_all, the histograms in ctx->ghist
int sort_segments_plan(o3r_ctx* ctx, const SortU32& sb, const uint32_t* seg_off, int n_seg, size_t per_seg_cap,
                       const GridParams* grids) {
    const size_t gh_bytes = (size_t)n_seg * kMaxPasses * kRsBins * 4;
    CU(ctx->ghist.ensure(gh_bytes));
    CU(ctx->plan_all.ensure((size_t)n_seg * sizeof(SortPlan)));
    ZERO(ctx->ghist.p, gh_bytes);
    SortPlan* plan = ctx->plan_all.as<SortPlan>();
    // digit layout from each segment's live index bits (<= 31 -> at most 4 passes)
    LAUNCH(k_rs_layout, cdiv(n_seg, 64), 64, 0, n_seg, grids, 31, plan);
    LAUNCH_N("k_rs_ghist_u32", (k_rs_ghist<uint32_t, 4>), dim3(std::max(1u, cdiv(per_seg_cap, kRsTile)), n_seg),
             kThreads, 0, sb.k0, seg_off, plan, ctx->ghist.as<uint32_t>());
    LAUNCH(k_rs_plan, n_seg, kThreads, 0, ctx->ghist.as<uint32_t>(), seg_off, plan, grids);
    return O3R_OK;
}

// sorts + reduces; out/out_off sized by the caller.  CNT_VOX receives the total.
int vg_sorted_reduce(o3r_ctx* ctx, const SortU32& sb, const float4* pts, const uint32_t* seg_off, int n_seg,
                     size_t per_seg_cap, const GridParams* grids, float ix, float iy, float iz, uint32_t min_points,
                     int z_shift, float4* out, uint32_t* out_off, uint64_t* out_keys, uint32_t* out_counts,
                     bool track_cells = false, const uint32_t* out_base = nullptr) {
    int rcs = sort_segments_plan(ctx, sb, seg_off, n_seg, per_seg_cap, grids);
    if (rcs) return rcs;
    SortPlan* plan = ctx->plan_all.as<SortPlan>();
    const bool fast = min_points <= 1 && !out_keys && !out_counts;
    int rc = sort_pairs<uint32_t>(ctx, sb.k0, sb.k1, sb.v0, sb.v1, seg_off, n_seg, per_seg_cap, plan, 4, 1,
                                  ctx->ghist.as<uint32_t>());
    if (rc) return rc;
    VgArgs A;
    A.keys0 = sb.k0; A.keys1 = sb.k1; A.vals0 = sb.v0; A.vals1 = sb.v1;
    A.seg_off = seg_off; A.plan = plan; A.grids = grids; A.pts = pts;
    A.tiles_ub = std::max(1u, cdiv(per_seg_cap, kTileV));
    A.min_points = min_points; A.z_shift = z_shift;
    A.lx_inv = ix; A.ly_inv = iy; A.lz_inv = iz;
    const size_t nt = (size_t)A.tiles_ub * n_seg;
    CU(ctx->head_cnt.ensure(nt * 4));
    CU(ctx->head_off.ensure(nt * 4));
    const dim3 grid(A.tiles_ub, n_seg);
    uint32_t* cnt = ctx->counters.as<uint32_t>();
    LAUNCH(k_vg_heads, grid, kThreads, 0, A, ctx->head_cnt.as<uint32_t>());
    LAUNCH(k_scan_u32, 1, kScanThreads, 0, ctx->head_cnt.as<uint32_t>(), ctx->head_off.as<uint32_t>(), (uint32_t)nt,
           cnt + CNT_VOX);
    if (fast) {
        const size_t wbytes = 64 + (nt + 1) * sizeof(RunCarry);
        CU(ctx->runwork.ensure(wbytes));
        ZERO(ctx->runwork.p, wbytes);
        LAUNCH(k_vg_reduce_w, (uint32_t)nt, kThreads, 0, A, ctx->head_off.as<uint32_t>(), cnt + CNT_VOX, out, out_off,
               n_seg, ctx->runwork.as<uint32_t>(), reinterpret_cast<RunCarry*>(ctx->runwork.as<char>() + 64),
               track_cells ? 1 : 0, ctx->inv_c, ctx->inv_cz, reinterpret_cast<int*>(cnt + CNT_CELLBB), out_base);
    } else
        LAUNCH(k_vg_reduce, grid, kThreads, 0, A, ctx->head_off.as<uint32_t>(), cnt + CNT_VOX, out, out_off, n_seg,
               out_keys, out_counts);
    return O3R_OK;
}

// ---- engine 2: merge `n` items (points or partial cells) into the resident accumulators --------------------------
inline int bits_for(long long range) { int b = 0; while ((1ll << b) <= range) ++b; return b; }

template <typename KeyT, typename Items>
int acc_build_cycle_t(o3r_ctx* ctx, Items items, size_t n, bool use_resident, const KeyCodec& kc, int total_bits) {
    const int passes = std::max(1, (total_bits + kRsMaxBits - 1) / kRsMaxBits);
    uint32_t* cnt = ctx->counters.as<uint32_t>();
    const size_t n4 = (n + 63) & ~(size_t)63;
    Tr tr("acc_build");
    CU(ctx->sortbuf.ensure(n4 * (2 * sizeof(KeyT) + 8)));
    tr.mark("sortbuf.ensure");
    KeyT* k0 = ctx->sortbuf.as<KeyT>();
    KeyT* k1 = k0 + n4;
    uint32_t* v0 = reinterpret_cast<uint32_t*>(k1 + n4);
    uint32_t* v1 = v0 + n4;
    CU(ctx->seg2.ensure(16));
    CU(ctx->ghist.ensure(kMaxPasses * kRsBins * 4));
    CU(ctx->plan_v2.ensure(sizeof(SortPlan)));
    const uint32_t seg_h[2] = {0u, (uint32_t)n};
    { int rcu = upload_small(ctx, ctx->seg2.p, seg_h, 8); if (rcu) return rcu; }
    ZERO(ctx->ghist.p, kMaxPasses * kRsBins * 4);
    ZERO(cnt + CNT_NEW, 4);
    const uint32_t* seg = ctx->seg2.as<uint32_t>();
    const uint32_t gk = std::min<uint32_t>(cdiv(n, kThreads), 148 * 8);
    SortPlan* plan = ctx->plan_v2.as<SortPlan>();
    LAUNCH(k_rs_layout, 1, 32, 0, 1, (const GridParams*)nullptr, total_bits, plan);
    LAUNCH_N("k_acc_key", (k_acc_key<KeyT, Items>), gk, kThreads, 0, items, (uint32_t)n, ctx->inv_c, ctx->inv_cz, kc, plan,
             k0, v0, ctx->ghist.as<uint32_t>());
    LAUNCH(k_rs_plan, 1, kThreads, 0, ctx->ghist.as<uint32_t>(), seg, plan, (const GridParams*)nullptr);
    tr.mark("key+plan");
    int rc = sort_pairs<KeyT>(ctx, k0, k1, v0, v1, seg, 1, n, plan, passes, 0, ctx->ghist.as<uint32_t>());
    if (rc) return rc;
    tr.mark("sort_pairs");
    AccArgs<KeyT> A;
    A.keys0 = k0; A.keys1 = k1; A.vals0 = v0; A.vals1 = v1;
    A.seg_off = seg; A.plan = plan;
    A.tiles_ub = cdiv(n, kTileV);
    A.kc = kc;
    const int cur = ctx->res_cur;
    A.res_keys = ctx->res_keys[cur].as<uint64_t>();
    A.res_acc = ctx->res_acc[cur].as<float4>();
    A.res_rgb = ctx->res_rgb[cur].as<uint4>();
    A.n_res_ptr = cnt + (use_resident ? CNT_NRES : CNT_ZERO);
    CU(ctx->head_cnt.ensure((size_t)A.tiles_ub * 4));
    CU(ctx->head_off.ensure((size_t)A.tiles_ub * 4));
    // one record per cell of the cycle (<= n_cyc_ub <= n), reserved in big steps so steady-state cycles never reallocate
    const size_t cyc_cap = std::max<size_t>(std::min<size_t>(n, ctx->n_cyc_ub ? ctx->n_cyc_ub : n), (size_t)1 << 20);
    CU(ctx->ckey.ensure(cyc_cap * 8));
    CU(ctx->cacc.ensure(cyc_cap * 16));
    CU(ctx->crgb.ensure(cyc_cap * 16));
    const size_t wbytes = 64 + ((size_t)A.tiles_ub + 1) * sizeof(RunCarry);
    CU(ctx->runwork.ensure(wbytes));
    ZERO(ctx->runwork.p, wbytes);
    tr.mark("ensures");
    LAUNCH_N("k_acc_heads", (k_acc_heads<KeyT>), A.tiles_ub, kThreads, 0, A, v0, v1, ctx->head_cnt.as<uint32_t>(),
             cnt + CNT_NEW);
    LAUNCH(k_scan_u32, 1, kScanThreads, 0, ctx->head_cnt.as<uint32_t>(), ctx->head_off.as<uint32_t>(), A.tiles_ub,
           cnt + CNT_CYC);
    LAUNCH_N("k_acc_reduce", (k_acc_reduce<KeyT, Items>), A.tiles_ub, kThreads, 0, A, items, ctx->head_off.as<uint32_t>(),
             ctx->ckey.as<uint64_t>(), ctx->cacc.as<float4>(), ctx->crgb.as<uint4>(), ctx->runwork.as<uint32_t>(),
             reinterpret_cast<RunCarry*>(ctx->runwork.as<char>() + 64));
    tr.mark("heads+reduce");
    return O3R_OK;
}

// Tightens the host's upper bound of the resident cell count from the read-back of the last merge.
int refresh_nres(o3r_ctx* ctx, bool block) {
    if (ctx->n_res_exact) return O3R_OK;
    if (block) CU(cudaEventSynchronize(ctx->ev_nres));
    else if (cudaEventQuery(ctx->ev_nres) != cudaSuccess) { cudaGetLastError(); return O3R_OK; }
    ctx->n_res_ub = *ctx->h_nres;
    ctx->n_res_exact = true;
    return O3R_OK;
}

// Builds the cycle's cell list (ckey/cacc/crgb, n_cyc) from items, continuing from the resident sums when
// `use_resident`.  `bb` = {imin,jmin,kmin,imax,jmax,kmax} of the items' combined-grid cells when the caller already
// knows it (host), else null.  Leaves h_counters[CNT_CYC], [CNT_NEW] valid.
template <typename Items>
int acc_build_cycle(o3r_ctx* ctx, const Items& items, size_t n, bool use_resident, const int* bb) {
    uint32_t* cnt = ctx->counters.as<uint32_t>();
    ctx->n_cyc = 0;
    ctx->n_cyc_ub = 0;
    ctx->h_counters[CNT_CYC] = ctx->h_counters[CNT_NEW] = 0;
    if (n == 0) return O3R_OK;
    { int rc = refresh_nres(ctx, false); if (rc) return rc; }
    if (n >= (1ull << 32)) return ctx->fail(O3R_ERR_INVALID, "batch too large for 32-bit indices");
    int hb[6];
    if (!bb) {
        LAUNCH(k_cellbb_init, 1, 32, 0, reinterpret_cast<int*>(cnt + CNT_CELLBB));
        LAUNCH_N("k_acc_cellbb", (k_acc_cellbb<Items>), std::min<uint32_t>(cdiv(n, kThreads), 148 * 8), kThreads, 0, items,
                 (uint32_t)n, ctx->inv_c, ctx->inv_cz, reinterpret_cast<int*>(cnt + CNT_CELLBB));
        int rc = read_counters(ctx);
        if (rc) return rc;
        memcpy(hb, ctx->h_counters + CNT_CELLBB, sizeof(hb));
        bb = hb;
    }
    const int lim = (1 << 20) - 1;
    for (int a = 0; a < 3; ++a)
        if (bb[a] < -lim || bb[3 + a] > lim || bb[3 + a] < bb[a])
            return ctx->fail(O3R_ERR_INVALID, "cloud extends beyond +-2^20 combined-grid cells (or is not finite)");
    KeyCodec kc;
    kc.imin = bb[0]; kc.jmin = bb[1]; kc.kmin = bb[2];
    kc.wi = bits_for((long long)bb[3] - bb[0]);
    kc.wj = bits_for((long long)bb[4] - bb[1]);
    const int total = kc.wi + kc.wj + bits_for((long long)bb[5] - bb[2]);
    // the cycle cannot touch more cells than its cell range holds (nor more than it has items)
    const double range_cells = ((double)bb[3] - bb[0] + 1) * ((double)bb[4] - bb[1] + 1) * ((double)bb[5] - bb[2] + 1);
    ctx->n_cyc_ub = (size_t)std::min((double)n, range_cells);
    if (total <= 32) return acc_build_cycle_t<uint32_t, Items>(ctx, items, n, use_resident, kc, total);
    return acc_build_cycle_t<uint64_t, Items>(ctx, items, n, use_resident, kc, total);
}

// Applies the cycle's cell list to the resident shard: in-place update of found cells, sorted insert of new ones.
// Entirely device-driven: the cycle's cell count, the number of new cells and the resident count are read from device
// memory, launches and buffers are sized by host-known upper bounds, and nothing here waits for the GPU.
int acc_apply_cycle(o3r_ctx* ctx) {
    const size_t n_cyc_ub = ctx->n_cyc_ub;
    if (n_cyc_ub == 0) return O3R_OK;
    uint32_t* cnt = ctx->counters.as<uint32_t>();
    const int cur = ctx->res_cur, nxt = cur ^ 1;
    const uint32_t tiles = cdiv(n_cyc_ub, kTileV);
    const size_t tot_ub = ctx->n_res_ub + n_cyc_ub;
    Tr tr("acc_apply");
    if (tot_ub >= (1ull << 32)) return ctx->fail(O3R_ERR_NOMEM, "resident shard could exceed 2^32 cells");
    CU(ctx->new_cnt.ensure((size_t)tiles * 4));
    CU(ctx->new_off.ensure((size_t)tiles * 4));
    // the shard grows a little every cycle: reserve in big steps (>= 4 M cells, doubling), a reallocation is a
    // cudaMalloc + cudaFree = a device-wide sync of several ms in the middle of the cycle
    if (ctx->res_keys[nxt].cap < tot_ub * 8 || ctx->res_acc[nxt].cap < tot_ub * 16 || ctx->res_rgb[nxt].cap < tot_ub * 16) {
        const size_t cells = std::max<size_t>(2 * tot_ub, (size_t)1 << 22);
        CU(ctx->res_keys[nxt].ensure(cells * 8));
        CU(ctx->res_acc[nxt].ensure(cells * 16));
        CU(ctx->res_rgb[nxt].ensure(cells * 16));
    }
    CU(ctx->new_keys.ensure(n_cyc_ub * 8));
    if (tr.on) fprintf(stderr, "[o3r trace] n_res_ub %zu n_cyc_ub %zu\n", ctx->n_res_ub, n_cyc_ub);
    tr.mark("ensures");
    LAUNCH(k_acc_update, tiles, kThreads, 0, cnt + CNT_CYC, ctx->cacc.as<float4>(), ctx->crgb.as<uint4>(),
           ctx->res_acc[cur].as<float4>(), ctx->res_rgb[cur].as<uint4>(), ctx->new_cnt.as<uint32_t>());
    LAUNCH(k_scan_u32, 1, kScanThreads, 0, ctx->new_cnt.as<uint32_t>(), ctx->new_off.as<uint32_t>(), tiles,
           cnt + CNT_NEWSCAN);
    LAUNCH(k_acc_place_new, tiles, kThreads, 0, cnt + CNT_CYC, ctx->ckey.as<uint64_t>(), ctx->cacc.as<float4>(),
           ctx->crgb.as<uint4>(), ctx->new_off.as<uint32_t>(), ctx->res_keys[cur].as<uint64_t>(), cnt + CNT_NRES,
           ctx->res_keys[nxt].as<uint64_t>(), ctx->res_acc[nxt].as<float4>(), ctx->res_rgb[nxt].as<uint4>(),
           ctx->new_keys.as<uint64_t>());
    if (ctx->n_res_ub) {
        const uint32_t g = std::min<uint32_t>(cdiv(ctx->n_res_ub, kThreads), 148 * 16);
        LAUNCH(k_acc_place_old, g, kThreads, 0, cnt + CNT_NRES, ctx->res_keys[cur].as<uint64_t>(),
               ctx->res_acc[cur].as<float4>(), ctx->res_rgb[cur].as<uint4>(), ctx->new_keys.as<uint64_t>(), cnt + CNT_NEW,
               ctx->res_keys[nxt].as<uint64_t>(), ctx->res_acc[nxt].as<float4>(), ctx->res_rgb[nxt].as<uint4>());
    }
    LAUNCH(k_acc_finish, 1, 32, 0, cnt + CNT_NRES, cnt + CNT_NEW);
    ctx->res_cur = nxt;
    ctx->n_res_ub = tot_ub;
    ctx->n_res_exact = false;
    CU(cudaMemcpyAsync(ctx->h_nres, cnt + CNT_NRES, 4, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaEventRecord(ctx->ev_nres, ctx->st));
    tr.mark("launches");
    return O3R_OK;
}

int acc_merge_points(o3r_ctx* ctx, const float4* pts, size_t n, const int* bb) {
    AccItemsPts items{pts, nullptr, nullptr};
    int rc = acc_build_cycle(ctx, items, n, true, bb);
    if (rc) return rc;
    return acc_apply_cycle(ctx);
}

int acc_merge_cells(o3r_ctx* ctx, const o3r_cell* cells, size_t n, const int* bb) {
    AccItemsCells items{cells, nullptr, nullptr};
    int rc = acc_build_cycle(ctx, items, n, true, bb);
    if (rc) return rc;
    return acc_apply_cycle(ctx);
}

int cloud_append_dev(o3r_ctx* ctx, const float4* pts, size_t n) {
    if (n == 0) return O3R_OK;
    CU(ctx->cloud.ensure((ctx->n_cloud + n) * 16, ctx->st, ctx->n_cloud * 16));
    CU(cudaMemcpyAsync(ctx->cloud.as<float4>() + ctx->n_cloud, pts, n * 16, cudaMemcpyDeviceToDevice, ctx->st));
    ctx->n_cloud += n;
    return O3R_OK;
}

// ---- pcl::StatisticalOutlierRemoval on the frames of a chunk (pose_functions.cpp:1673-1686) -----------------------------
// in: pts with segment offsets seg_off[n_seg + 1] (device).  out: ctx->sor_pts / ctx->sor_off (same layout, kept points
// in their original order).  Uses the sort buffers (free again afterwards) and ctx->bbox / ctx->spts as scratch.
int sor_filter(o3r_ctx* ctx, const SortU32& sb, const float4* pts, const uint32_t* seg_off, int n_seg, size_t per_seg_cap,
               int mean_k, double stddev_mul) {
    if (mean_k < 1) return ctx->fail(O3R_ERR_INVALID, "sor_mean_k must be positive");
    if (mean_k + 1 > kSorMaxK) return ctx->fail(O3R_ERR_INVALID, "sor_mean_k too large (max 127)");
    const size_t cap = per_seg_cap * n_seg;
    const uint32_t tiles = std::max(1u, cdiv(per_seg_cap, kTileV));
    CU(ctx->sor_pts.ensure(cap * 16));
    CU(ctx->sor_off.ensure((size_t)(n_seg + 1) * 4));
    CU(ctx->sor_dist.ensure(cap * 4));
    CU(ctx->sor_skeys.ensure(cap * 4));
    CU(ctx->sor_svals.ensure(cap * 4));
    CU(ctx->spts.ensure(cap * 16));
    CU(ctx->sor_grids.ensure((size_t)n_seg * sizeof(SorGrid)));
    CU(ctx->sor_pgrids.ensure((size_t)n_seg * sizeof(GridParams)));
    CU(ctx->sor_rows.ensure((size_t)n_seg * kSorRowsCap * 8));
    CU(ctx->sor_thr.ensure((size_t)n_seg * 8));
    CU(ctx->sor_cnt.ensure((size_t)tiles * n_seg * 4));
    CU(ctx->sor_cntoff.ensure((size_t)tiles * n_seg * 4));
    CU(ctx->bbox.ensure((size_t)n_seg * 6 * 4));
    SorGrid* grids = ctx->sor_grids.as<SorGrid>();
    GridParams* pgrids = ctx->sor_pgrids.as<GridParams>();
    uint32_t* rowb = ctx->sor_rows.as<uint32_t>();
    uint32_t* rowe = rowb + (size_t)n_seg * kSorRowsCap;
    uint32_t* cnt = ctx->counters.as<uint32_t>();
    const uint32_t gl = std::min<uint32_t>(std::max(1u, cdiv(per_seg_cap, kThreads)), 148 * 4);
    LAUNCH(k_bbox_init, cdiv((size_t)n_seg * 6, kThreads), kThreads, 0, ctx->bbox.as<uint32_t>(), n_seg);
    LAUNCH(k_bbox_pts, dim3(tiles, n_seg), kThreads, 0, pts, seg_off, 0, ctx->bbox.as<uint32_t>());
    LAUNCH(k_sor_grid, cdiv(n_seg, 64), 64, 0, n_seg, ctx->bbox.as<uint32_t>(), seg_off, mean_k, grids, pgrids);
    const size_t heap_bytes = (size_t)(mean_k + 1) * kSorThreads * 4;
    CU(cudaFuncSetAttribute(k_sor_knn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heap_bytes));
    CU(cudaFuncSetAttribute(k_sor_calib, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heap_bytes));
    SortPlan* plan = ctx->plan_all.as<SortPlan>();
    // grid build, twice: on the density guess, then on the cell the calibration derives from 128 exact sample queries per frame
    for (int round = 0; round < 2; ++round) {
        LAUNCH(k_sor_key, dim3(gl, n_seg), kThreads, 0, pts, seg_off, grids, sb.k0);
        int rc = sort_segments_plan(ctx, sb, seg_off, n_seg, per_seg_cap, pgrids);
        if (rc) return rc;
        plan = ctx->plan_all.as<SortPlan>();
        rc = sort_pairs<uint32_t>(ctx, sb.k0, sb.k1, sb.v0, sb.v1, seg_off, n_seg, per_seg_cap, plan, 4, 1, ctx->ghist.as<uint32_t>());
        if (rc) return rc;
        LAUNCH(k_sor_rows_clear, dim3(std::min<uint32_t>(cdiv(kSorRowsCap, kThreads), 256), n_seg), kThreads, 0, grids, rowb, rowe);
        LAUNCH(k_sor_rows, dim3(gl, n_seg), kThreads, 0, sb.k0, sb.k1, sb.v0, sb.v1, plan, pts, seg_off, grids, rowb, rowe,
               ctx->sor_skeys.as<uint32_t>(), ctx->sor_svals.as<uint32_t>(), ctx->spts.as<float4>());
        if (round == 0)
            LAUNCH(k_sor_calib, n_seg, kSorThreads, heap_bytes, ctx->spts.as<float4>(), ctx->sor_skeys.as<uint32_t>(), seg_off, grids,
                   pgrids, ctx->bbox.as<uint32_t>(), rowb, rowe, mean_k);
    }
    CU(ctx->sor_hard.ensure(cap * 8 + 64));
    uint32_t* n_hard = reinterpret_cast<uint32_t*>(ctx->sor_hard.as<char>() + cap * 8);
    ZERO(n_hard, 4);
    CU(cudaFuncSetAttribute(k_sor_knn_hard, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heap_bytes));
    LAUNCH(k_sor_knn, dim3(std::max(1u, cdiv(per_seg_cap, kSorThreads)), n_seg), kSorThreads, heap_bytes, ctx->spts.as<float4>(),
           ctx->sor_skeys.as<uint32_t>(), ctx->sor_svals.as<uint32_t>(), seg_off, grids, rowb, rowe, mean_k,
           ctx->sor_dist.as<float>(), ctx->sor_hard.as<uint2>(), n_hard);
    LAUNCH(k_sor_knn_hard, std::max(1u, cdiv(cap, kSorThreads)), kSorThreads, heap_bytes, ctx->spts.as<float4>(),
           ctx->sor_skeys.as<uint32_t>(), ctx->sor_svals.as<uint32_t>(), seg_off, grids, rowb, rowe, mean_k,
           ctx->sor_dist.as<float>(), ctx->sor_hard.as<uint2>(), n_hard);
    LAUNCH(k_sor_stats, n_seg, kThreads, 0, ctx->sor_dist.as<float>(), seg_off, stddev_mul, ctx->sor_thr.as<double>());
    LAUNCH(k_sor_count, dim3(tiles, n_seg), kThreads, 0, ctx->sor_dist.as<float>(), seg_off, ctx->sor_thr.as<double>(), tiles,
           ctx->sor_cnt.as<uint32_t>());
    LAUNCH(k_scan_u32, 1, kScanThreads, 0, ctx->sor_cnt.as<uint32_t>(), ctx->sor_cntoff.as<uint32_t>(), (uint32_t)((size_t)tiles * n_seg),
           cnt + CNT_PTS);
    LAUNCH(k_sor_compact, dim3(tiles, n_seg), kThreads, 0, pts, ctx->sor_dist.as<float>(), seg_off, ctx->sor_thr.as<double>(), tiles,
           ctx->sor_cntoff.as<uint32_t>(), cnt + CNT_PTS, n_seg, ctx->sor_pts.as<float4>(), ctx->sor_off.as<uint32_t>());
    return O3R_OK;
}

// ---- the batched per-frame path ------------------------------------------------------------------------------------
struct BatchOpts {
    bool merge = true;          // append / merge into the resident cloud
    uint8_t* mask_dev = nullptr;  // optional validity mask output
    bool mask_only = false;
};

// stage A for one chunk of frames (descriptors at `fr`).  V1 mode: points + leaf indices into the chunk scratch.
// dont_downsample mode: points straight into the batch buffer at the device-side running offset.
template <int DT>
int launch_stage_a(o3r_ctx* ctx, const AParams& P, const FrameDev* fr, int n, const BatchOpts& opt, float4* pts_out,
                   uint32_t* keys_out, const uint32_t* out_base, uint32_t* goff) {
    const dim3 grid(P.tiles_per_frame, n);
    uint32_t* cnt = ctx->counters.as<uint32_t>();
    LAUNCH(k_bbox_init, cdiv((size_t)n * 6, kThreads), kThreads, 0, ctx->bbox.as<uint32_t>(), n);
    LAUNCH_N("k_pre", (k_pre<DT>), grid, kThreads, 0, P, fr, ctx->tile_cnt.as<uint32_t>(), ctx->bbox.as<uint32_t>(),
             opt.mask_dev);
    if (opt.mask_only) return O3R_OK;
    LAUNCH(k_scan_u32, 1, kScanThreads, 0, ctx->tile_cnt.as<uint32_t>(), ctx->tile_off.as<uint32_t>(),
           (uint32_t)((size_t)P.tiles_per_frame * n), cnt + CNT_PTS);
    LAUNCH(k_a_post, cdiv(n + 1, kThreads), kThreads, 0, n, P.tiles_per_frame, ctx->tile_off.as<uint32_t>(),
           cnt + CNT_PTS, ctx->bbox.as<uint32_t>(), ctx->inv_f, ctx->inv_f, ctx->inv_f, P.want_keys,
           ctx->frame_off.as<uint32_t>(), ctx->grids.as<GridParams>(), out_base, goff);
    LAUNCH_N("k_emit", (k_emit<DT>), grid, kThreads, 0, P, fr, ctx->tile_off.as<uint32_t>(),
             ctx->frame_off.as<uint32_t>(), ctx->grids.as<GridParams>(), pts_out, keys_out, out_base);
    return O3R_OK;
}

// ---- input staging (host entry points) --------------------------------------------------------------------------------
struct StageGeom { size_t es, dstep, dplane, cstep, cplane, lstep, lplane; };

StageGeom stage_geom(const o3r_params& p, int disp_type) {
    StageGeom G;
    G.es = disp_elem(disp_type);
    G.dstep = (((size_t)p.cols * G.es) + 15) & ~(size_t)15; G.dplane = G.dstep * p.rows;
    G.cstep = (((size_t)p.cols * 3) + 15) & ~(size_t)15;     G.cplane = G.cstep * p.rows;
    G.lstep = ((size_t)p.cols + 15) & ~(size_t)15;           G.lplane = G.lstep * p.rows;
    return G;
}

void frames_signature(const o3r_frame* frames, int n, std::vector<const void*>& sig) {
    sig.clear();
    sig.reserve((size_t)n * 4);
    for (int i = 0; i < n; ++i) {
        sig.push_back(frames[i].disp); sig.push_back(frames[i].bgr);
        sig.push_back(frames[i].labels); sig.push_back(frames[i].kp_xy);
    }
}

// sizes staging set `set` for the frames and fills their device-side descriptors (everything but T)
int stage_layout(o3r_ctx* ctx, const o3r_frame* frames, int n, bool label_mode, int J, const StageGeom& G, int set,
                 std::vector<FrameDev>& fd) {
    o3r_ctx::Staging& S = ctx->stg[set];
    size_t kp_total = 0, coef_total = 0;
    for (int i = 0; i < n; ++i) {
        const o3r_frame& f = frames[i];
        kp_total += (J != 1 && f.kp_xy && f.n_kp > 0) ? (size_t)f.n_kp * 2 : 0;
        coef_total += label_mode ? (size_t)std::max(f.n_planes, 0) * 3 : 0;
    }
    if (!label_mode) CU(S.disp.ensure(G.dplane * n));
    CU(S.bgr.ensure(G.cplane * n));
    if (label_mode) { CU(S.labels.ensure(G.lplane * n)); CU(S.coef.ensure(std::max<size_t>(coef_total, 1) * 8)); }
    if (kp_total) CU(S.kp.ensure(kp_total * 4));
    size_t kp_at = 0, coef_at = 0;
    for (int i = 0; i < n; ++i) {
        const o3r_frame& f = frames[i];
        FrameDev& d = fd[i];
        const int nk = (J != 1 && f.kp_xy && f.n_kp > 0) ? f.n_kp : 0;
        d.disp = label_mode ? nullptr : S.disp.as<uint8_t>() + G.dplane * i; d.disp_step = G.dstep;
        d.bgr = S.bgr.as<uint8_t>() + G.cplane * i; d.bgr_step = G.cstep;
        d.labels = label_mode ? S.labels.as<uint8_t>() + G.lplane * i : nullptr; d.labels_step = G.lstep;
        d.plane_coef = label_mode ? S.coef.as<double>() + coef_at : nullptr;
        d.kp_xy = nk ? S.kp.as<float>() + kp_at : nullptr;
        d.n_planes = label_mode ? std::max(f.n_planes, 0) : 0;
        d.n_kp = nk;
        kp_at += (size_t)nk * 2; coef_at += (size_t)d.n_planes * 3;
    }
    return O3R_OK;
}

// issues the H2D copies of frames [f0, f0 + nc): frames alternate between the two copy streams, and the first stream
// then waits for the second, so an event recorded on st_copy after this call covers every copy
int stage_copy(o3r_ctx* ctx, const o3r_frame* frames, const std::vector<FrameDev>& fd, int f0, int nc, bool label_mode,
               const StageGeom& G) {
    const o3r_params& p = ctx->p;
    // Only the pixels the path can read cross PCIe: the scan ROI x in [x0, cols-bb), y in [bb, rows-bb)
    // (pose_functions.cpp:1062,1094-1095; keypoints outside it are rejected too), widened by the blur window's reach for
    // the disparity plane.  The device planes keep the full-image layout, so the kernels index as before.
    const int halo = p.blur_kernel > 1 ? p.blur_kernel / 2 + 1 : 0;
    const int cx0 = std::min(p.cols, std::max(0, p.cols_start_aft_cutout)), cx1 = std::max(cx0, p.cols - p.bounding_box);
    const int cy0 = std::min(p.rows, std::max(0, p.bounding_box)), cy1 = std::max(cy0, p.rows - p.bounding_box);
    const int dx0 = std::max(0, cx0 - halo), dx1 = std::min(p.cols, cx1 + halo);
    const int dy0 = std::max(0, cy0 - halo), dy1 = std::min(p.rows, cy1 + halo);
    if (cx1 <= cx0 || cy1 <= cy0) return O3R_OK;   // empty ROI: nothing is ever read
    for (int i = f0; i < f0 + nc; ++i) {
        const o3r_frame& f = frames[i];
        const FrameDev& d = fd[i];
        cudaStream_t cs = (i & 1) ? ctx->st_copy2 : ctx->st_copy;
        if (!label_mode) {
            CU(cudaMemcpy2DAsync((uint8_t*)d.disp + (size_t)dy0 * G.dstep + (size_t)dx0 * G.es, G.dstep,
                                 (const uint8_t*)f.disp + (size_t)dy0 * f.disp_step + (size_t)dx0 * G.es, f.disp_step,
                                 (size_t)(dx1 - dx0) * G.es, dy1 - dy0, cudaMemcpyHostToDevice, cs));
        } else {
            CU(cudaMemcpy2DAsync((uint8_t*)d.labels + (size_t)cy0 * G.lstep + cx0, G.lstep,
                                 f.labels + (size_t)cy0 * f.labels_step + cx0, f.labels_step, (size_t)(cx1 - cx0), cy1 - cy0,
                                 cudaMemcpyHostToDevice, cs));
            if (d.n_planes)
                CU(cudaMemcpyAsync((void*)d.plane_coef, f.plane_coef, (size_t)d.n_planes * 24, cudaMemcpyHostToDevice, cs));
        }
        CU(cudaMemcpy2DAsync((uint8_t*)d.bgr + (size_t)cy0 * G.cstep + (size_t)cx0 * 3, G.cstep,
                             f.bgr + (size_t)cy0 * f.bgr_step + (size_t)cx0 * 3, f.bgr_step, (size_t)(cx1 - cx0) * 3, cy1 - cy0,
                             cudaMemcpyHostToDevice, cs));
        if (d.n_kp) CU(cudaMemcpyAsync((void*)d.kp_xy, f.kp_xy, (size_t)d.n_kp * 8, cudaMemcpyHostToDevice, cs));
    }
    if (nc > 1) {
        CU(cudaEventRecord(ctx->ev_copy2, ctx->st_copy2));
        CU(cudaStreamWaitEvent(ctx->st_copy, ctx->ev_copy2, 0));
    }
    return O3R_OK;
}

// issues the copies of a recorded prefetch request into a free staging set
int flush_deferred_prefetch(o3r_ctx* ctx) {
    if (!ctx->deferred.pending) return O3R_OK;
    ctx->deferred.pending = false;
    const o3r_params& p = ctx->p;
    const o3r_frame* frames = ctx->deferred.frames.data();
    const int n = (int)ctx->deferred.frames.size(), disp_type = ctx->deferred.disp_type;
    const bool label_mode = p.use_segment_labels && frames[0].labels && frames[0].plane_coef;
    const StageGeom G = stage_geom(p, disp_type);
    const int set = ctx->free_stage_set();
    o3r_ctx::Prefetch& pf = ctx->prefetch[set];
    pf.valid = false;
    std::vector<FrameDev> fd(n);
    int rc = stage_layout(ctx, frames, n, label_mode, p.jump_pixels, G, set, fd);
    if (rc) return rc;
    rc = stage_copy(ctx, frames, fd, 0, n, label_mode, G);
    if (rc) return rc;
    if (!pf.ev) CU(cudaEventCreateWithFlags(&pf.ev, cudaEventDisableTiming));
    CU(cudaEventRecord(pf.ev, ctx->st_copy));
    frames_signature(frames, n, pf.sig);
    pf.n = n; pf.disp_type = disp_type; pf.seq = ++ctx->prefetch_seq; pf.valid = true;
    return O3R_OK;
}

// The batched per-frame path.  `frames` hold host pointers (host_inputs: every frame is copied to device staging
// on the copy stream, chunk by chunk, overlapping the previous chunk's kernels) or device pointers.
int frames_cloud_impl(o3r_ctx* ctx, const o3r_frame* frames, int n, int disp_type, uint32_t* frame_counts,
                      const BatchOpts& opt, bool host_inputs) {
    const o3r_params& p = ctx->p;
    static const bool trace = getenv("O3R_TRACE") != nullptr;
    ctx->busy_set = -1;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tr0 = now();
    if (n <= 0) { ctx->last_n = 0; ctx->last_total = 0; return O3R_OK; }
    if (n > 65535) return ctx->fail(O3R_ERR_INVALID, "too many frames in one batch");
    if (disp_type < 0 || disp_type > 3) return ctx->fail(O3R_ERR_INVALID, "bad disp_type");
    const bool label_mode = p.use_segment_labels && frames[0].labels && frames[0].plane_coef;
    if (p.use_segment_labels && !label_mode && disp_type != O3R_DISP_F64)
        return ctx->fail(O3R_ERR_INVALID, "use_segment_labels needs an F64 disparity plane or labels + plane_coef");
    const bool blur = p.blur_kernel > 1;
    if (blur) {  // cv::bilateralFilter / medianBlur reject CV_64F: the reference throws and yields an empty cloud
        if (disp_type != O3R_DISP_U8 || p.use_segment_labels)
            return ctx->fail(O3R_ERR_INVALID, "blur_kernel > 1 requires u8 disparity (OpenCV rejects CV_64F)");
        if (p.blur_mode == O3R_BLUR_MEDIAN && (p.blur_kernel & 1) == 0)
            return ctx->fail(O3R_ERR_INVALID, "median blur needs an odd blur_kernel (cv::medianBlur asserts)");
        if (p.blur_mode != O3R_BLUR_MEDIAN && p.blur_mode != O3R_BLUR_BOX && p.blur_mode != O3R_BLUR_BILATERAL)
            return ctx->fail(O3R_ERR_INVALID, "unknown blur_mode");
        if (p.blur_kernel > kBlurMaxK) return ctx->fail(O3R_ERR_INVALID, "blur_kernel too large (max 127)");
    }
    const int J = p.jump_pixels;
    int max_kp = 0;
    size_t kp_total = 0, coef_total = 0;
    for (int i = 0; i < n; ++i) {
        const o3r_frame& f = frames[i];
        if (!f.bgr || (!f.disp && !label_mode)) return ctx->fail(O3R_ERR_INVALID, "frame without disparity/colour");
        if (label_mode && (!f.labels || !f.plane_coef)) return ctx->fail(O3R_ERR_INVALID, "frame without labels");
        const int nk = (J != 1 && f.kp_xy && f.n_kp > 0) ? f.n_kp : 0;
        max_kp = std::max(max_kp, nk);
        kp_total += (size_t)nk * 2;
        coef_total += label_mode ? (size_t)std::max(f.n_planes, 0) * 3 : 0;
    }

    AParams P;
    memset(&P, 0, sizeof(P));
    P.rows = p.rows; P.cols = p.cols; P.x0 = p.cols_start_aft_cutout; P.bb = p.bounding_box; P.J = J;
    P.nx = ctx->nx; P.ny = ctx->ny; P.npix = ctx->npix;
    P.kp_tiles = (int)cdiv((size_t)max_kp, kTileA);
    P.tiles_per_frame = P.kp_tiles + (int)cdiv(P.npix, kTileA);
    P.label_mode = label_mode;
    P.canon = ctx->canon;
    P.use_lut = ctx->canon && disp_type == O3R_DISP_U8 && !label_mode;
    P.thr_i = (int)std::floor(p.min_disparity);
    P.min_disp = p.min_disparity; P.div = p.disp_divisor;
    for (int i = 0; i < 16; ++i) P.q[i] = p.Q[i];
    P.lut_r = ctx->lut_r.as<double>(); P.lut_z = ctx->lut_z.as<float>();
    P.want_keys = !p.dont_downsample && !opt.mask_only;
    P.want_bbox = P.want_keys;
    if (P.tiles_per_frame == 0) {  // nothing to scan (J == 0 and no keypoints)
        ctx->last_n = n; ctx->last_total = 0; ctx->last_off.assign(n + 1, 0);
        if (frame_counts) std::fill(frame_counts, frame_counts + n, 0u);
        return O3R_OK;
    }

    // ---- device copies of the inputs (host entry points) and frame descriptors
    const size_t es = disp_elem(disp_type);
    const StageGeom G = stage_geom(p, disp_type);
    const size_t blur_step = G.lstep;
    // inputs already on their way (o3r_frames_prefetch of exactly these frames)?
    bool prefetched = false;
    int set = 0;
    if (host_inputs) {
        std::vector<const void*> sig;
        frames_signature(frames, n, sig);
        // the OLDEST matching prefetch: a caller that recycles its host buffers issues the next cycle's prefetch
        // (same pointers) before this call, and that copy is still in flight
        for (int s = 0; s < 2; ++s) {
            const o3r_ctx::Prefetch& pf = ctx->prefetch[s];
            if (pf.valid && pf.n == n && pf.disp_type == disp_type && pf.sig == sig &&
                (!prefetched || pf.seq < ctx->prefetch[set].seq)) {
                prefetched = true; set = s;
            }
        }
        if (!prefetched && ctx->deferred.pending && (int)ctx->deferred.frames.size() == n &&
            ctx->deferred.disp_type == disp_type) {   // recorded but not issued yet (nothing ran in between): issue it now
            std::vector<const void*> dsig;
            frames_signature(ctx->deferred.frames.data(), n, dsig);
            if (dsig == sig) {
                int rcf = flush_deferred_prefetch(ctx);
                if (rcf) return rcf;
                for (int s = 0; s < 2; ++s)
                    if (ctx->prefetch[s].valid && ctx->prefetch[s].seq == ctx->prefetch_seq) { prefetched = true; set = s; }
            }
        }
        if (prefetched) ctx->prefetch[set].valid = false;
        if (!prefetched) set = ctx->free_stage_set();
        ctx->busy_set = set;
    }
    if (blur) CU(ctx->d_blur.ensure((size_t)n * p.rows * blur_step));
    std::vector<FrameDev> fd(n);
    std::vector<BlurJob> jobs(blur ? n : 0);
    bool vec = (J == 1) && (ctx->nx % 4 == 0) && (P.x0 % 4 == 0) && !label_mode;
    std::vector<FrameDev> fd_stage;   // where the H2D copies land (fd[i].disp is redirected to the blurred plane below)
    if (host_inputs) {
        int rc = stage_layout(ctx, frames, n, label_mode, J, G, set, fd);
        if (rc) return rc;
        if (!prefetched) fd_stage = fd;
    }
    for (int i = 0; i < n; ++i) {
        const o3r_frame& f = frames[i];
        FrameDev& d = fd[i];
        if (!host_inputs) {
            const int nk = (J != 1 && f.kp_xy && f.n_kp > 0) ? f.n_kp : 0;
            d.disp = (const uint8_t*)f.disp; d.disp_step = f.disp_step;
            d.bgr = f.bgr; d.bgr_step = f.bgr_step;
            d.labels = label_mode ? f.labels : nullptr; d.labels_step = f.labels_step;
            d.plane_coef = label_mode ? f.plane_coef : nullptr;
            d.kp_xy = nk ? f.kp_xy : nullptr;
            d.n_planes = label_mode ? std::max(f.n_planes, 0) : 0;
            d.n_kp = nk;
        }
        for (int k = 0; k < 12; ++k) d.T[k] = f.T[k];
        if (blur) {
            jobs[i].src = d.disp; jobs[i].sstep = d.disp_step;
            jobs[i].dst = ctx->d_blur.as<uint8_t>() + (size_t)i * p.rows * blur_step; jobs[i].dstep = blur_step;
            d.disp = jobs[i].dst; d.disp_step = blur_step;
        }
        if (!label_mode) vec = vec && ((uintptr_t)d.disp % 16 == 0) && (d.disp_step % (4 * es) == 0);
        vec = vec && ((uintptr_t)d.bgr % 4 == 0) && (d.bgr_step % 4 == 0);
    }
    P.vec = vec;
    CU(ctx->d_frames.ensure((size_t)n * sizeof(FrameDev)));
    { int rcu = upload_small(ctx, ctx->d_frames.p, fd.data(), (size_t)n * sizeof(FrameDev)); if (rcu) return rcu; }
    if (blur) {
        CU(ctx->d_blurjobs.ensure((size_t)n * sizeof(BlurJob)));
        { int rcu = upload_small(ctx, ctx->d_blurjobs.p, jobs.data(), (size_t)n * sizeof(BlurJob)); if (rcu) return rcu; }
    }

    // ---- buffers: batch-wide outputs, chunk-sized scratch
    // host inputs: chunks let the copies overlap the kernels; device inputs: one launch sequence for the whole batch
    const int chunk = std::max(1, std::min(n, (host_inputs && !prefetched) ? ctx->chunk_frames : ctx->chunk_frames_dev));
    if (prefetched && trace) {
        const bool done = cudaEventQuery(ctx->prefetch[set].ev) == cudaSuccess;
        cudaGetLastError();
        fprintf(stderr, "[o3r trace] prefetched inputs %s at entry\n", done ? "ready" : "still in flight");
    }
    if (prefetched) CU(cudaStreamWaitEvent(ctx->st, ctx->prefetch[set].ev, 0));
    const size_t per_frame_cap = (size_t)P.npix + (size_t)max_kp;
    const size_t cap_batch = per_frame_cap * n, cap_chunk = per_frame_cap * chunk;
    if (cap_batch >= (1ull << 32)) return ctx->fail(O3R_ERR_INVALID, "batch exceeds 2^32 samples; use fewer frames");
    const size_t n_tiles = (size_t)P.tiles_per_frame * chunk;
    CU(ctx->tile_cnt.ensure(n_tiles * 4));
    CU(ctx->tile_off.ensure(n_tiles * 4));
    CU(ctx->bbox.ensure((size_t)chunk * 6 * 4));
    CU(ctx->frame_off.ensure((size_t)(chunk + 1) * 4));
    CU(ctx->grids.ensure((size_t)chunk * sizeof(GridParams)));
    SortU32 sb{nullptr, nullptr, nullptr, nullptr};
    uint32_t* cnt = ctx->counters.as<uint32_t>();
    if (!opt.mask_only) {
        CU(ctx->vox_off.ensure((size_t)(n + 1) * 4));
        if (P.want_keys) {
            CU(ctx->pts.ensure(cap_chunk * 16));
            CU(ctx->vox.ensure(cap_batch * 16));
            int rc = carve_sort_u32(ctx, cap_chunk, sb);
            if (rc) return rc;
            if (!ctx->retain()) LAUNCH(k_cellbb_init, 1, 32, 0, reinterpret_cast<int*>(cnt + CNT_CELLBB));
            ctx->tiled_now = ctx->tiled() && !(ctx->tiled_poor && (ctx->tiled_batches % 16) != 0);
            ++ctx->tiled_batches;
            if (ctx->tiled_now) {   // worst case one partial per item; never reached in practice (~1/8)
                CU(ctx->partials.ensure(cap_batch * sizeof(o3r_cell)));
                CU(ctx->pr_status.ensure(((size_t)cdiv(cap_chunk, kPrTile) + 16) * 4));
                ZERO(cnt + CNT_PART, 8);
            }
        } else {
            CU(ctx->pts.ensure(cap_batch * 16));
        }
        ZERO(cnt + CNT_BASE, 4);
        ZERO(ctx->vox_off.p, 4);
    }
    ctx->last_is_vox = P.want_keys;
    uint32_t* goff = ctx->vox_off.as<uint32_t>();

    for (int f0 = 0; f0 < n; f0 += chunk) {
        const int nc = std::min(chunk, n - f0);
        if (host_inputs && !prefetched) {  // H2D of this chunk on the copy stream; the compute stream waits for its event only
            int rc = stage_copy(ctx, frames, fd_stage, f0, nc, label_mode, G);
            if (rc) return rc;
            cudaEvent_t ev = ctx->chunk_event(f0 / chunk);
            CU(cudaEventRecord(ev, ctx->st_copy));
            CU(cudaStreamWaitEvent(ctx->st, ev, 0));
        }
        const FrameDev* fr = ctx->d_frames.as<FrameDev>() + f0;
        if (blur) {
            const int rx0 = P.x0, rx1 = p.cols - p.bounding_box, ry0 = p.bounding_box, ry1 = p.rows - p.bounding_box;
            if (rx1 > rx0 && ry1 > ry0) {
                const dim3 g(cdiv(rx1 - rx0, kBlurStrip), cdiv(ry1 - ry0, kBlurRows), nc);
                const size_t sm = blur_smem(p.blur_kernel, p.blur_mode);
                const BlurJob* bj = ctx->d_blurjobs.as<BlurJob>() + f0;
                if (p.blur_mode == O3R_BLUR_BILATERAL) {
                    BilateralLut L;
                    int rcl = bilateral_lut(ctx, p.blur_kernel, &L);
                    if (rcl) return rcl;
                    const dim3 gb(cdiv(rx1 - rx0, kBilTX), cdiv(ry1 - ry0, kBilTY), nc);
                    LAUNCH(k_bilateral, gb, dim3(kBilTX, kBilTY), bilateral_smem(L.radius, L.maxk), bj, L, p.rows, p.cols, rx0, ry0, rx1, ry1);
                } else if (p.blur_mode == O3R_BLUR_MEDIAN)
                    LAUNCH_N("k_blur_median", (k_blur<O3R_BLUR_MEDIAN>), g, kBlurRows, sm, bj, p.rows, p.cols, p.blur_kernel, rx0, ry0, rx1, ry1);
                else
                    LAUNCH_N("k_blur_box", (k_blur<O3R_BLUR_BOX>), g, kBlurRows, sm, bj, p.rows, p.cols, p.blur_kernel, rx0, ry0, rx1, ry1);
            }
        }
        // stage A: V1 mode writes the chunk scratch; dont_downsample mode writes the batch buffer at the running base
        float4* pts_out = ctx->pts.as<float4>();
        const uint32_t* base_a = P.want_keys ? nullptr : cnt + CNT_BASE;
        uint32_t* goff_a = P.want_keys ? nullptr : goff + f0;
        int rc;
        switch (disp_type) {
            case O3R_DISP_U8: rc = launch_stage_a<O3R_DISP_U8>(ctx, P, fr, nc, opt, pts_out, sb.k0, base_a, goff_a); break;
            case O3R_DISP_U16: rc = launch_stage_a<O3R_DISP_U16>(ctx, P, fr, nc, opt, pts_out, sb.k0, base_a, goff_a); break;
            case O3R_DISP_F32: rc = launch_stage_a<O3R_DISP_F32>(ctx, P, fr, nc, opt, pts_out, sb.k0, base_a, goff_a); break;
            default: rc = launch_stage_a<O3R_DISP_F64>(ctx, P, fr, nc, opt, pts_out, sb.k0, base_a, goff_a); break;
        }
        if (rc) return rc;
        if (opt.mask_only) return O3R_OK;   // (single frame)
        const float4* vg_pts = ctx->pts.as<float4>();
        const uint32_t* vg_off = ctx->frame_off.as<uint32_t>();
        if (P.want_keys && p.sor_mean_k > 0 && J > 0) {
            // StatisticalOutlierRemoval first (pose_functions.cpp:1673-1686); the VoxelGrid then sees the filtered cloud, so
            // its bbox, grid and leaf indices are recomputed from the kept points
            rc = sor_filter(ctx, sb, vg_pts, vg_off, nc, per_frame_cap, p.sor_mean_k, p.sor_stddev_mul);
            if (rc) return rc;
            vg_pts = ctx->sor_pts.as<float4>();
            vg_off = ctx->sor_off.as<uint32_t>();
            const uint32_t tl = std::max(1u, cdiv(per_frame_cap, kTileV));
            LAUNCH(k_bbox_init, cdiv((size_t)nc * 6, kThreads), kThreads, 0, ctx->bbox.as<uint32_t>(), nc);
            LAUNCH(k_bbox_pts, dim3(tl, nc), kThreads, 0, vg_pts, vg_off, 0, ctx->bbox.as<uint32_t>());
            LAUNCH(k_grid_params, cdiv(nc, 64), 64, 0, nc, ctx->bbox.as<uint32_t>(), ctx->inv_f, ctx->inv_f, ctx->inv_f,
                   ctx->grids.as<GridParams>());
            LAUNCH(k_vg_key, dim3(std::min<uint32_t>(std::max(1u, cdiv(per_frame_cap, kThreads)), 148 * 4), nc), kThreads, 0, vg_pts,
                   vg_off, ctx->grids.as<GridParams>(), 0, sb.k0);
        }
        if (P.want_keys) {  // per-frame VoxelGrid, leaf voxel_size / 5 (pose_functions.cpp:1698), appended at the base
            rc = vg_sorted_reduce(ctx, sb, vg_pts, vg_off, nc, per_frame_cap,
                                  ctx->grids.as<GridParams>(), ctx->inv_f, ctx->inv_f, ctx->inv_f, 0, 0,
                                  ctx->vox.as<float4>(), goff + f0, nullptr, nullptr, !ctx->retain(), cnt + CNT_BASE);
            if (rc) return rc;
            if (ctx->tiled_now) {   // group the chunk's voxel centroids by combined-grid cell, tile by tile
                const uint32_t pt = cdiv(cap_chunk, kPrTile);
                uint32_t* stw = ctx->pr_status.as<uint32_t>();
                ZERO(stw, ((size_t)pt + 16) * 4);
                ZERO(cnt + CNT_PARTCHUNK, 4);
                LAUNCH(k_cell_prereduce, pt, kThreads, 0, ctx->vox.as<float4>(), cnt + CNT_BASE, cnt + CNT_VOX, ctx->inv_c,
                       ctx->inv_cz, ctx->partials.as<o3r_cell>(), cnt + CNT_PART, cnt + CNT_PARTCHUNK, stw, stw + pt);
                LAUNCH(k_add_u32, 1, 32, 0, cnt + CNT_PART, cnt + CNT_PARTCHUNK);
            }
            LAUNCH(k_add_base, 1, 32, 0, cnt + CNT_BASE, cnt + CNT_VOX, goff + f0 + nc);
        } else {
            LAUNCH(k_add_base, 1, 32, 0, cnt + CNT_BASE, cnt + CNT_PTS, goff + f0 + nc);
        }
    }

    // ---- per-frame output offsets (+ the combined-grid cell range) back to the host: the one sync of the frame path
    if (ctx->h_offs_cap < (size_t)n + 1) {
        if (ctx->h_offs) cudaFreeHost(ctx->h_offs);
        ctx->h_offs_cap = std::max<size_t>(n + 1, 256);
        CU(cudaMallocHost((void**)&ctx->h_offs, ctx->h_offs_cap * 4));
    }
    CU(cudaMemcpyAsync(ctx->h_offs, goff, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, ctx->st));
    ctx->last_has_cellbb = ctx->last_is_vox && !ctx->retain();
    if (ctx->last_has_cellbb)
        CU(cudaMemcpyAsync(ctx->h_counters + CNT_CELLBB, cnt + CNT_CELLBB, 24, cudaMemcpyDeviceToHost, ctx->st));
    // the kernels of this call are queued: now issue the copies of a recorded prefetch (next cycle's inputs).  The
    // staging set they go to is not the one these kernels read.
    if (ctx->deferred.pending && !opt.mask_only) {
        int rcf = flush_deferred_prefetch(ctx);
        if (rcf) return rcf;
    }
    ctx->last_has_partials = ctx->last_is_vox && ctx->tiled() && ctx->tiled_now;
    if (ctx->last_has_partials)
        CU(cudaMemcpyAsync(ctx->h_counters + CNT_PART, cnt + CNT_PART, 4, cudaMemcpyDeviceToHost, ctx->st));
    const double tr1 = now();
    CU(cudaStreamSynchronize(ctx->st));
    ctx->busy_set = -1;
    ctx->last_partials = ctx->last_has_partials ? ctx->h_counters[CNT_PART] : 0;
    // The tile pre-reduction pays only when it reduces: a 40-byte partial replaces a 16-byte voxel in the merge.  On grids
    // finer than the point spacing (e.g. 4K at voxel_size 0.01) nearly every voxel is its own cell: merge the voxels then.
    if (ctx->last_has_partials) {
        ctx->tiled_poor = ctx->last_partials * 2 > ctx->h_offs[n];
        if (ctx->tiled_poor) { ctx->last_has_partials = false; ctx->last_partials = 0; }
    }
    const double tr2 = now();
    if (ctx->last_has_cellbb) memcpy(ctx->last_cellbb, ctx->h_counters + CNT_CELLBB, 24);
    ctx->last_off.assign(ctx->h_offs, ctx->h_offs + n + 1);
    ctx->last_n = n;
    ctx->last_total = ctx->last_off[n];
    if (frame_counts)
        for (int i = 0; i < n; ++i) frame_counts[i] = ctx->last_off[i + 1] - ctx->last_off[i];
    if (!opt.merge || ctx->defer_merge) return O3R_OK;
    const float4* outp = ctx->last_is_vox ? ctx->vox.as<float4>() : ctx->pts.as<float4>();
    if (ctx->retain()) return cloud_append_dev(ctx, outp, ctx->last_total);
    const int* bbp = ctx->last_has_cellbb ? ctx->last_cellbb : nullptr;
    const int rcm = ctx->last_has_partials ? acc_merge_cells(ctx, ctx->partials.as<o3r_cell>(), ctx->last_partials, bbp)
                                           : acc_merge_points(ctx, outp, ctx->last_total, bbp);
    if (trace) fprintf(stderr, "[o3r trace] frames_cloud: enqueue %.3f ms, sync wait %.3f ms, merge enqueue %.3f ms\n", tr1 - tr0, tr2 - tr1, now() - tr2);
    return rcm;
}

int copy_out(o3r_ctx* ctx, const float4* src, size_t n, o3r_point* out, size_t cap, size_t* n_out) {
    if (n_out) *n_out = n;
    if (!out && cap == 0) return O3R_OK;
    if (cap < n) return ctx->fail(O3R_ERR_CAPACITY, "output buffer too small");
    if (n) {
        CU(cudaMemcpyAsync(out, src, n * 16, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
    }
    return O3R_OK;
}

}  // namespace

// =====================================================================================================================
extern "C" {

int o3r_version(void) { return O3R_VERSION; }

const char* o3r_last_error(const o3r_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int o3r_create(const o3r_params* params, o3r_ctx** out_ctx) {
    if (!params || !out_ctx) { g_create_err = "null argument"; return O3R_ERR_INVALID; }
    *out_ctx = nullptr;
    const o3r_params& p = *params;
    if (p.rows <= 0 || p.cols <= 0 || p.bounding_box < 0 || p.cols_start_aft_cutout < 0 || p.jump_pixels < 0 ||
        !(p.voxel_size > 0) || p.max_batch_frames < 1) {
        g_create_err = "invalid o3r_params";
        return O3R_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= p.device) {
        g_create_err = std::string("no usable CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
        return O3R_ERR_CUDA;
    }
    e = cudaSetDevice(p.device);
    if (e != cudaSuccess) { g_create_err = cudaGetErrorString(e); return O3R_ERR_CUDA; }
    o3r_ctx* ctx = new o3r_ctx();
    ctx->p = p;
    const int J = p.jump_pixels;
    if (J > 0) {  // pose_functions.cpp:1094-1128
        ctx->ny = std::max(0, (p.rows - 2 * p.bounding_box + J - 1) / J);
        ctx->nx = std::max(0, (p.cols - p.bounding_box - p.cols_start_aft_cutout + J - 1) / J);
    }
    ctx->npix = (uint32_t)ctx->nx * (uint32_t)ctx->ny;
    const double* Q = p.Q;
    ctx->canon = Q[1] == 0 && Q[2] == 0 && Q[4] == 0 && Q[6] == 0 && Q[8] == 0 && Q[9] == 0 && Q[10] == 0 &&
                 Q[12] == 0 && Q[13] == 0;
    ctx->leaf_f = (float)(p.voxel_size / 5);          // pose_functions.cpp:1698
    ctx->inv_f = 1.0f / ctx->leaf_f;                  // PCL: inverse_leaf_size_ = Ones / leaf_size_
    ctx->leaf_c = (float)p.voxel_size;                // pose_functions.cpp:1694
    ctx->inv_c = 1.0f / ctx->leaf_c;
    ctx->inv_cz = 1.0f / 1000.0f;
    auto bail = [&](cudaError_t er, const char* what) {
        g_create_err = std::string(what) + ": " + cudaGetErrorString(er);
        o3r_destroy(ctx);
        return O3R_ERR_CUDA;
    };
    if ((e = cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    if ((e = cudaStreamCreateWithFlags(&ctx->st_copy, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    if ((e = cudaStreamCreateWithFlags(&ctx->st_copy2, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_copy2, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    if (const char* cf = getenv("O3R_CHUNK_FRAMES")) ctx->chunk_frames = std::max(1, atoi(cf));
    if (const char* cf = getenv("O3R_CHUNK_FRAMES_DEV")) ctx->chunk_frames_dev = std::max(1, atoi(cf));
    if ((e = cudaMallocHost((void**)&ctx->h_counters, CNT_N * 4)) != cudaSuccess) return bail(e, "pinned");
    if ((e = cudaMallocHost((void**)&ctx->h_nres, 4)) != cudaSuccess) return bail(e, "pinned");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_nres, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    if ((e = ctx->counters.ensure(CNT_N * 4)) != cudaSuccess) return bail(e, "counters");
    if ((e = cudaMemsetAsync(ctx->counters.p, 0, CNT_N * 4, ctx->st)) != cudaSuccess) return bail(e, "memset");
    // 1/(q14*d+q15) and (float)(q11 * that) for u8 disparities
    double lr[256];
    float lz[256];
    for (int d = 0; d < 256; ++d) {
        const double v3 = Q[14] * (double)d + Q[15];
        lr[d] = 1.0 / v3;
        lz[d] = (float)(Q[11] * lr[d]);
    }
    if ((e = ctx->lut_r.ensure(sizeof(lr))) != cudaSuccess) return bail(e, "lut");
    if ((e = ctx->lut_z.ensure(sizeof(lz))) != cudaSuccess) return bail(e, "lut");
    cudaMemcpyAsync(ctx->lut_r.p, lr, sizeof(lr), cudaMemcpyHostToDevice, ctx->st);
    cudaMemcpyAsync(ctx->lut_z.p, lz, sizeof(lz), cudaMemcpyHostToDevice, ctx->st);
    if ((e = cudaStreamSynchronize(ctx->st)) != cudaSuccess) return bail(e, "init sync");
    e = cudaFuncSetAttribute(k_rs_onesweep<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)rs_scatter_smem<uint64_t>());
    if (e != cudaSuccess) return bail(e, "smem attr (is this an sm_100a device?)");
    e = cudaFuncSetAttribute(k_rs_onesweep<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)rs_scatter_smem<uint32_t>());
    if (e != cudaSuccess) return bail(e, "smem attr");
    // shared-memory carve-out per kernel (percent of the 228 KB array; the rest is L1).  Measured on B200:
    // the run-reduce kernels gather points, so they want L1 as much as occupancy.
    {
        auto env_or = [](const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; };
        const int cv_vg = env_or("O3R_CARVEOUT_VG", 50), cv_acc = env_or("O3R_CARVEOUT_ACC", 50),
                  cv_sort = env_or("O3R_CARVEOUT_SORT", 100), cv_emit = env_or("O3R_CARVEOUT_EMIT", -1);
        cudaFuncSetAttribute(k_vg_reduce_w, cudaFuncAttributePreferredSharedMemoryCarveout, cv_vg);
        cudaFuncSetAttribute(k_acc_reduce<uint32_t, AccItemsPts>, cudaFuncAttributePreferredSharedMemoryCarveout, cv_acc);
        cudaFuncSetAttribute(k_acc_reduce<uint64_t, AccItemsPts>, cudaFuncAttributePreferredSharedMemoryCarveout, cv_acc);
        cudaFuncSetAttribute(k_acc_reduce<uint32_t, AccItemsCells>, cudaFuncAttributePreferredSharedMemoryCarveout, cv_acc);
        cudaFuncSetAttribute(k_acc_reduce<uint64_t, AccItemsCells>, cudaFuncAttributePreferredSharedMemoryCarveout, cv_acc);
        cudaFuncSetAttribute(k_rs_onesweep<uint32_t>, cudaFuncAttributePreferredSharedMemoryCarveout, cv_sort);
        cudaFuncSetAttribute(k_rs_onesweep<uint64_t>, cudaFuncAttributePreferredSharedMemoryCarveout, cv_sort);
        if (cv_emit >= 0) {
            cudaFuncSetAttribute(k_emit<O3R_DISP_U8>, cudaFuncAttributePreferredSharedMemoryCarveout, cv_emit);
            cudaFuncSetAttribute(k_emit<O3R_DISP_U16>, cudaFuncAttributePreferredSharedMemoryCarveout, cv_emit);
            cudaFuncSetAttribute(k_emit<O3R_DISP_F32>, cudaFuncAttributePreferredSharedMemoryCarveout, cv_emit);
            cudaFuncSetAttribute(k_emit<O3R_DISP_F64>, cudaFuncAttributePreferredSharedMemoryCarveout, cv_emit);
        }
    }
    const int bs = (int)blur_smem(kBlurMaxK, O3R_BLUR_MEDIAN);
    cudaFuncSetAttribute(k_blur<O3R_BLUR_MEDIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, bs);
    cudaFuncSetAttribute(k_blur<O3R_BLUR_BOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, bs);
    *out_ctx = ctx;
    return O3R_OK;
}

void o3r_destroy(o3r_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->p.device);
    if (ctx->st_copy2) cudaStreamSynchronize(ctx->st_copy2);
    if (ctx->st_copy) cudaStreamSynchronize(ctx->st_copy);
    if (ctx->st) cudaStreamSynchronize(ctx->st);
    for (auto& pf : ctx->prefetch) if (pf.ev) cudaEventDestroy(pf.ev);
    DevBuf* bufs[] = {&ctx->lut_r, &ctx->lut_z, &ctx->d_disp, &ctx->stg[0].disp, &ctx->stg[0].bgr, &ctx->stg[0].labels,
                      &ctx->stg[0].coef, &ctx->stg[0].kp, &ctx->stg[1].disp, &ctx->stg[1].bgr, &ctx->stg[1].labels,
                      &ctx->stg[1].coef, &ctx->stg[1].kp,
                      &ctx->d_frames, &ctx->d_blur, &ctx->d_blurjobs, &ctx->bil_lut, &ctx->pp_labels, &ctx->pp_disp, &ctx->pp_sums,
                      &ctx->pp_rows, &ctx->pp_coef, &ctx->tile_cnt, &ctx->tile_off, &ctx->bbox,
                      &ctx->frame_off, &ctx->grids, &ctx->counters, &ctx->pts, &ctx->sortbuf, &ctx->hist,
                      &ctx->plan_all, &ctx->plan_v2, &ctx->ghist, &ctx->head_cnt, &ctx->head_off, &ctx->vox,
                      &ctx->vox_off, &ctx->seg2, &ctx->tmat, &ctx->mask, &ctx->runwork, &ctx->spts, &ctx->res_keys[0], &ctx->res_keys[1],
                      &ctx->res_acc[0], &ctx->res_acc[1], &ctx->res_rgb[0], &ctx->res_rgb[1], &ctx->ckey, &ctx->cacc,
                      &ctx->crgb, &ctx->partials, &ctx->pr_status, &ctx->sor_hard, &ctx->sor_pts, &ctx->sor_off, &ctx->sor_dist, &ctx->sor_grids,
                      &ctx->sor_pgrids, &ctx->sor_rows, &ctx->sor_thr, &ctx->sor_skeys, &ctx->sor_svals, &ctx->sor_cnt, &ctx->sor_cntoff, &ctx->new_cnt, &ctx->new_off, &ctx->new_keys, &ctx->okeys, &ctx->cloud};
    for (DevBuf* b : bufs) b->release();
    for (auto& r : ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_nres) cudaFreeHost(ctx->h_nres);
    if (ctx->ev_nres) cudaEventDestroy(ctx->ev_nres);
    if (ctx->h_offs) cudaFreeHost(ctx->h_offs);
    for (auto e : ctx->chunk_ev) cudaEventDestroy(e);
    if (ctx->ev_copy2) cudaEventDestroy(ctx->ev_copy2);
    if (ctx->st_copy2) cudaStreamDestroy(ctx->st_copy2);
    if (ctx->st_copy) cudaStreamDestroy(ctx->st_copy);
    if (ctx->st) cudaStreamDestroy(ctx->st);
    delete ctx;
}

uint64_t o3r_launch_count(const o3r_ctx* ctx) { return ctx ? ctx->launches : 0; }
size_t o3r_last_batch_partials(const o3r_ctx* ctx) { return (ctx && ctx->last_has_partials) ? ctx->last_partials : 0; }
void* o3r_stream(o3r_ctx* ctx) { return ctx ? (void*)ctx->st : nullptr; }

int o3r_sync(o3r_ctx* ctx) {
    if (!ctx) return O3R_ERR_INVALID;
    CU(cudaStreamSynchronize(ctx->st));
    return O3R_OK;
}

int o3r_profile(o3r_ctx* ctx, int enable) {
    if (!ctx) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaStreamSynchronize(ctx->st));
    for (auto& r : ctx->prof) { ctx->ev_pool.push_back(r.a); ctx->ev_pool.push_back(r.b); }
    ctx->prof.clear();
    ctx->profiling = enable != 0;
    return O3R_OK;
}

int o3r_profile_read(o3r_ctx* ctx, char* buf, size_t cap) {
    if (!ctx || !buf || cap == 0) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaStreamSynchronize(ctx->st));
    struct Agg { std::string name; uint64_t n; double ms; };
    std::vector<Agg> agg;
    for (auto& r : ctx->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        std::string nm = r.name;
        size_t i = 0;
        for (; i < agg.size(); ++i) if (agg[i].name == nm) break;
        if (i == agg.size()) agg.push_back({nm, 0, 0.0});
        agg[i].n += 1; agg[i].ms += ms;
    }
    std::string out;
    char line[256];
    for (auto& a : agg) {
        snprintf(line, sizeof(line), "%s\t%llu\t%.6f\n", a.name.c_str(), (unsigned long long)a.n, a.ms);
        out += line;
    }
    if (out.size() + 1 > cap) return ctx->fail(O3R_ERR_CAPACITY, "profile buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return O3R_OK;
}

int o3r_set_defer_merge(o3r_ctx* ctx, int defer) {
    if (!ctx) return O3R_ERR_INVALID;
    ctx->defer_merge = defer != 0;
    return O3R_OK;
}

void* o3r_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void o3r_host_free(void* p) { if (p) cudaFreeHost(p); }

int o3r_frames_cloud_dev(o3r_ctx* ctx, const o3r_frame* frames, int n, int disp_type, uint32_t* frame_counts) {
    if (!ctx || (n > 0 && !frames)) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    return frames_cloud_impl(ctx, frames, n, disp_type, frame_counts, BatchOpts(), false);
}

int o3r_frames_cloud(o3r_ctx* ctx, const o3r_frame* frames, int n, int disp_type, uint32_t* frame_counts) {
    if (!ctx || (n > 0 && !frames)) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    return frames_cloud_impl(ctx, frames, n, disp_type, frame_counts, BatchOpts(), true);
}

int o3r_frames_prefetch(o3r_ctx* ctx, const o3r_frame* frames, int n, int disp_type) {
    if (!ctx || n <= 0 || !frames) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    if (disp_type < 0 || disp_type > 3) return ctx->fail(O3R_ERR_INVALID, "bad disp_type");
    const o3r_params& p = ctx->p;
    const bool label_mode = p.use_segment_labels && frames[0].labels && frames[0].plane_coef;
    for (int i = 0; i < n; ++i)
        if (!frames[i].bgr || (!frames[i].disp && !label_mode) || (label_mode && (!frames[i].labels || !frames[i].plane_coef)))
            return ctx->fail(O3R_ERR_INVALID, "frame without disparity/colour");
    int rc = flush_deferred_prefetch(ctx);   // at most one request waits; an older one goes out now
    if (rc) return rc;
    ctx->deferred.frames.assign(frames, frames + n);
    ctx->deferred.disp_type = disp_type;
    ctx->deferred.pending = true;
    return O3R_OK;
}

int o3r_frame_cloud(o3r_ctx* ctx, const o3r_frame* frame, int disp_type, o3r_point* out, size_t cap, size_t* n_out) {
    if (!ctx || !frame) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    if (n_out) *n_out = 0;
    BatchOpts opt;
    opt.merge = false;
    int rc = frames_cloud_impl(ctx, frame, 1, disp_type, nullptr, opt, true);
    if (rc) return rc;
    const float4* src = ctx->last_is_vox ? ctx->vox.as<float4>() : ctx->pts.as<float4>();
    return copy_out(ctx, src, ctx->last_total, out, cap, n_out);
}

int o3r_frame_mask(o3r_ctx* ctx, const o3r_frame* frame, int disp_type, uint8_t* mask, size_t cap, size_t* n_scanned) {
    if (!ctx || !frame) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    if (n_scanned) *n_scanned = ctx->npix;
    if (!mask && cap == 0) return O3R_OK;
    if (cap < ctx->npix) return ctx->fail(O3R_ERR_CAPACITY, "mask buffer too small");
    if (ctx->npix == 0) return O3R_OK;
    CU(ctx->mask.ensure(ctx->npix));
    BatchOpts opt;
    opt.merge = false; opt.mask_only = true; opt.mask_dev = ctx->mask.as<uint8_t>();
    int rc = frames_cloud_impl(ctx, frame, 1, disp_type, nullptr, opt, true);
    if (rc) return rc;
    CU(cudaMemcpyAsync(mask, ctx->mask.p, ctx->npix, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return O3R_OK;
}

int o3r_last_batch_points(o3r_ctx* ctx, o3r_point* out, size_t cap, size_t* n_out) {
    if (!ctx) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    const float4* src = ctx->last_is_vox ? ctx->vox.as<float4>() : ctx->pts.as<float4>();
    return copy_out(ctx, src, ctx->last_total, out, cap, n_out);
}

int o3r_cloud_transform(o3r_ctx* ctx, const float T[16]) {
    if (!ctx || !T) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    if (!ctx->retain())
        return ctx->fail(O3R_ERR_UNSUPPORTED,
                         "o3r_cloud_transform needs O3R_MERGE_RETAIN: accumulated cells cannot be re-binned");
    if (ctx->n_cloud == 0) return O3R_OK;
    CU(ctx->tmat.ensure(64));
    { int rcu = upload_small(ctx, ctx->tmat.p, T, 48); if (rcu) return rcu; }
    const uint32_t g = std::min<uint32_t>(cdiv(ctx->n_cloud, kThreads), 148 * 16);
    LAUNCH(k_transform_pts, g, kThreads, 0, ctx->cloud.as<float4>(), ctx->n_cloud, ctx->tmat.as<float>());
    return O3R_OK;
}

int o3r_cloud_append(o3r_ctx* ctx, const o3r_point* pts, size_t n) {
    if (!ctx || (n && !pts)) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    if (n == 0) return O3R_OK;
    if (ctx->retain()) {
        CU(ctx->cloud.ensure((ctx->n_cloud + n) * 16, ctx->st, ctx->n_cloud * 16));
        CU(cudaMemcpyAsync(ctx->cloud.as<float4>() + ctx->n_cloud, pts, n * 16, cudaMemcpyHostToDevice, ctx->st));
        ctx->n_cloud += n;
        return O3R_OK;
    }
    CU(ctx->vox.ensure(n * 16));
    CU(cudaMemcpyAsync(ctx->vox.p, pts, n * 16, cudaMemcpyHostToDevice, ctx->st));
    ctx->last_n = 0; ctx->last_total = 0; ctx->last_has_cellbb = false;   // the batch buffer was overwritten
    ctx->last_has_partials = false; ctx->last_partials = 0;
    return acc_merge_points(ctx, ctx->vox.as<float4>(), n, nullptr);
}

int o3r_cloud_size(o3r_ctx* ctx, size_t* n) {
    if (!ctx || !n) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->retain()) { *n = ctx->n_cloud; return O3R_OK; }
    int rc = refresh_nres(ctx, true);
    if (rc) return rc;
    *n = ctx->n_res_ub;
    return O3R_OK;
}

int o3r_cloud_clear(o3r_ctx* ctx) {
    if (!ctx) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    CU(cudaStreamSynchronize(ctx->st));
    ctx->n_cloud = 0; ctx->n_res_ub = 0; ctx->n_res_exact = true;
    ZERO(ctx->counters.as<uint32_t>() + CNT_NRES, 4);
    return O3R_OK;
}

// engine 1 on an arbitrary device cloud of one segment
static int voxel_grid_dev(o3r_ctx* ctx, const float4* pts, size_t n, float ix, float iy, float iz, uint32_t min_points,
                          int z_shift, float4* out, uint64_t* out_keys, uint32_t* out_counts, size_t* n_out,
                          int* passthrough) {
    if (n >= (1ull << 32)) return ctx->fail(O3R_ERR_INVALID, "cloud too large for 32-bit indices");
    uint32_t* cnt = ctx->counters.as<uint32_t>();
    CU(ctx->seg2.ensure(16));
    const uint32_t seg_h[2] = {0u, (uint32_t)n};
    { int rcu = upload_small(ctx, ctx->seg2.p, seg_h, 8); if (rcu) return rcu; }
    const uint32_t* seg = ctx->seg2.as<uint32_t>();
    CU(ctx->bbox.ensure(6 * 4));
    CU(ctx->grids.ensure(sizeof(GridParams)));
    CU(ctx->vox_off.ensure(2 * 4));
    SortU32 sb;
    int rc = carve_sort_u32(ctx, n, sb);
    if (rc) return rc;
    const uint32_t tiles = cdiv(n, kTileV);
    LAUNCH(k_bbox_init, 1, kThreads, 0, ctx->bbox.as<uint32_t>(), 1);
    LAUNCH(k_bbox_pts, dim3(tiles, 1), kThreads, 0, pts, seg, z_shift, ctx->bbox.as<uint32_t>());
    LAUNCH(k_grid_params, 1, 32, 0, 1, ctx->bbox.as<uint32_t>(), ix, iy, iz, ctx->grids.as<GridParams>());
    LAUNCH(k_vg_key, dim3(std::min<uint32_t>(cdiv(n, kThreads), 148 * 16), 1), kThreads, 0, pts, seg,
           ctx->grids.as<GridParams>(), z_shift, sb.k0);
    rc = vg_sorted_reduce(ctx, sb, pts, seg, 1, n, ctx->grids.as<GridParams>(), ix, iy, iz, min_points, z_shift, out,
                          ctx->vox_off.as<uint32_t>(), out_keys, out_counts);
    if (rc) return rc;
    GridParams g;
    CU(cudaMemcpyAsync(&g, ctx->grids.p, sizeof(g), cudaMemcpyDeviceToHost, ctx->st));
    rc = read_counters(ctx);
    if (rc) return rc;
    (void)cnt;
    *n_out = ctx->h_counters[CNT_VOX];
    if (passthrough) *passthrough = g.passthrough;
    return O3R_OK;
}

// downsamplePtCloud(cloud_big, true) into a device buffer: *res = result, *m = records
// When *exact_pending comes back true, *m is only an upper bound and the exact record count is still on the device
// (cnt[CNT_EMIT]): the caller queues its read-back (and the output copy) behind the kernels and syncs ONCE.
static int cloud_downsample_dev(o3r_ctx* ctx, const float4** res, size_t* m, bool* exact_pending) {
    *res = nullptr; *m = 0;
    if (exact_pending) *exact_pending = false;
    if (ctx->p.dont_downsample) {  // pose.cpp:533-536: cloud_small = cloud_big
        *res = ctx->cloud.as<float4>(); *m = ctx->n_cloud;
        return O3R_OK;
    }
    uint32_t* cnt = ctx->counters.as<uint32_t>();
    if (ctx->retain()) {  // one-shot pcl::VoxelGrid over cloud_big, exactly pose_functions.cpp:1654-1709
        if (ctx->n_cloud == 0) return O3R_OK;
        CU(ctx->ckey.ensure(ctx->n_cloud * 16));   // result points (<= n_cloud)
        int rc = voxel_grid_dev(ctx, ctx->cloud.as<float4>(), ctx->n_cloud, ctx->inv_c, ctx->inv_c, ctx->inv_cz,
                                ctx->p.min_points_per_voxel, 1, ctx->ckey.as<float4>(), nullptr, nullptr, m, nullptr);
        if (rc) return rc;
        *res = ctx->ckey.as<float4>();
        return O3R_OK;
    }
    // device-driven: launches are sized by the host's upper bound of the resident count, the kernels read the exact one
    { int rc0 = refresh_nres(ctx, false); if (rc0) return rc0; }
    const size_t n_ub = ctx->n_res_ub;
    if (n_ub == 0) return O3R_OK;
    const int cur = ctx->res_cur;
    const uint32_t tiles = cdiv(n_ub, kTileV);
    CU(ctx->new_cnt.ensure((size_t)tiles * 4));
    CU(ctx->new_off.ensure((size_t)tiles * 4));
    CU(ctx->cacc.ensure(n_ub * 16));
    LAUNCH(k_acc_emit_cnt, tiles, kThreads, 0, cnt + CNT_NRES, ctx->res_acc[cur].as<float4>(), ctx->p.min_points_per_voxel,
           ctx->new_cnt.as<uint32_t>());
    LAUNCH(k_scan_u32, 1, kScanThreads, 0, ctx->new_cnt.as<uint32_t>(), ctx->new_off.as<uint32_t>(), tiles, cnt + CNT_EMIT);
    LAUNCH(k_acc_emit, tiles, kThreads, 0, cnt + CNT_NRES, ctx->res_acc[cur].as<float4>(), ctx->res_rgb[cur].as<uint4>(),
           ctx->p.min_points_per_voxel, ctx->new_off.as<uint32_t>(), ctx->cacc.as<float4>());
    *res = ctx->cacc.as<float4>();
    *m = n_ub;   // upper bound; the exact count is cnt[CNT_EMIT] once the stream has run (callers read it back)
    if (exact_pending) *exact_pending = true;
    return O3R_OK;
}

// one read-back of the exact output count (and of the resident count, which tightens the host's bound)
static int downsample_finish(o3r_ctx* ctx, size_t* m) {
    uint32_t* cnt = ctx->counters.as<uint32_t>();
    CU(cudaMemcpyAsync(ctx->h_counters, cnt, CNT_N * 4, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    *m = ctx->h_counters[CNT_EMIT];
    ctx->n_res_ub = ctx->h_counters[CNT_NRES];
    ctx->n_res_exact = true;
    return O3R_OK;
}

int o3r_cloud_downsample(o3r_ctx* ctx, o3r_point* out, size_t cap, size_t* n_out) {
    if (!ctx) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    if (n_out) *n_out = 0;
    const float4* res;
    size_t m;
    bool pending;
    int rc = cloud_downsample_dev(ctx, &res, &m, &pending);
    if (rc) return rc;
    if (!pending) return copy_out(ctx, res, m, out, cap, n_out);
    // accumulators: the copy of up to min(cap, bound) records is queued behind the emit kernels, one sync for everything
    const size_t ncopy = (out && cap) ? std::min(cap, m) : 0;
    if (ncopy) CU(cudaMemcpyAsync(out, res, ncopy * 16, cudaMemcpyDeviceToHost, ctx->st));
    rc = downsample_finish(ctx, &m);
    if (rc) return rc;
    if (n_out) *n_out = m;
    if (!out && cap == 0) return O3R_OK;
    if (cap < m) return ctx->fail(O3R_ERR_CAPACITY, "output buffer too small");
    return O3R_OK;
}

int o3r_cloud_downsample_dev(o3r_ctx* ctx, const o3r_point** dev_out, size_t* n_out) {
    if (!ctx || !n_out) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    const float4* res;
    bool pending;
    int rc = cloud_downsample_dev(ctx, &res, n_out, &pending);
    if (rc) return rc;
    if (pending) { rc = downsample_finish(ctx, n_out); if (rc) return rc; }
    else CU(cudaStreamSynchronize(ctx->st));
    if (dev_out) *dev_out = reinterpret_cast<const o3r_point*>(res);
    return O3R_OK;
}

int o3r_voxel_grid(o3r_ctx* ctx, const o3r_point* pts, size_t n, float lx, float ly, float lz, unsigned min_points,
                   o3r_point* out, size_t cap, size_t* n_out, uint64_t* keys, uint32_t* counts, int* passthrough) {
    if (!ctx || !n_out || (n && !pts)) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    *n_out = 0;
    if (passthrough) *passthrough = 0;
    if (n == 0) return O3R_OK;
    if (!(lx > 0) || !(ly > 0) || !(lz > 0)) return ctx->fail(O3R_ERR_INVALID, "leaf size must be positive");
    CU(ctx->pts.ensure(n * 16));
    CU(ctx->vox.ensure(n * 16));
    CU(ctx->ckey.ensure(n * 8));
    CU(ctx->new_keys.ensure(n * 4));
    ctx->last_n = 0; ctx->last_total = 0;
    CU(cudaMemcpyAsync(ctx->pts.p, pts, n * 16, cudaMemcpyHostToDevice, ctx->st));
    size_t m = 0;
    int rc = voxel_grid_dev(ctx, ctx->pts.as<float4>(), n, 1.0f / lx, 1.0f / ly, 1.0f / lz, min_points, 0,
                            ctx->vox.as<float4>(), ctx->ckey.as<uint64_t>(), ctx->new_keys.as<uint32_t>(), &m, passthrough);
    if (rc) return rc;
    *n_out = m;
    if (!out && cap == 0) return O3R_OK;
    if (cap < m) return ctx->fail(O3R_ERR_CAPACITY, "output buffer too small");
    if (m) {
        CU(cudaMemcpyAsync(out, ctx->vox.p, m * 16, cudaMemcpyDeviceToHost, ctx->st));
        if (keys) CU(cudaMemcpyAsync(keys, ctx->ckey.p, m * 8, cudaMemcpyDeviceToHost, ctx->st));
        if (counts) CU(cudaMemcpyAsync(counts, ctx->new_keys.p, m * 4, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
    }
    return O3R_OK;
}

int o3r_blur_u8(o3r_ctx* ctx, const uint8_t* src, size_t src_step, int rows, int cols, int kernel, int mode,
                uint8_t* dst, size_t dst_step) {
    if (!ctx || !src || !dst || rows <= 0 || cols <= 0) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    if (kernel < 1 || kernel > kBlurMaxK) return ctx->fail(O3R_ERR_INVALID, "kernel must be in 1..127");
    if (mode == O3R_BLUR_MEDIAN && (kernel & 1) == 0)
        return ctx->fail(O3R_ERR_INVALID, "median blur needs an odd kernel (cv::medianBlur asserts)");
    if (mode != O3R_BLUR_MEDIAN && mode != O3R_BLUR_BOX && mode != O3R_BLUR_BILATERAL)
        return ctx->fail(O3R_ERR_INVALID, "unknown blur mode");
    const size_t step = ((size_t)cols + 15) & ~(size_t)15, plane = step * rows;
    CU(ctx->d_disp.ensure(plane));
    CU(ctx->d_blur.ensure(plane));
    CU(ctx->d_blurjobs.ensure(sizeof(BlurJob)));
    CU(cudaMemcpy2DAsync(ctx->d_disp.p, step, src, src_step, cols, rows, cudaMemcpyHostToDevice, ctx->st));
    BlurJob job{ctx->d_disp.as<uint8_t>(), step, ctx->d_blur.as<uint8_t>(), step};
    CU(cudaMemcpyAsync(ctx->d_blurjobs.p, &job, sizeof(job), cudaMemcpyHostToDevice, ctx->st));
    const dim3 g(cdiv(cols, kBlurStrip), cdiv(rows, kBlurRows), 1);
    const size_t sm = blur_smem(kernel, mode);
    if (mode == O3R_BLUR_BILATERAL) {
        BilateralLut L;
        int rcl = bilateral_lut(ctx, kernel, &L);
        if (rcl) return rcl;
        const dim3 gb(cdiv(cols, kBilTX), cdiv(rows, kBilTY), 1);
        LAUNCH(k_bilateral, gb, dim3(kBilTX, kBilTY), bilateral_smem(L.radius, L.maxk), ctx->d_blurjobs.as<BlurJob>(), L, rows, cols, 0, 0, cols, rows);
    } else if (mode == O3R_BLUR_MEDIAN)
        LAUNCH_N("k_blur_median", (k_blur<O3R_BLUR_MEDIAN>), g, kBlurRows, sm, ctx->d_blurjobs.as<BlurJob>(), rows, cols, kernel, 0, 0, cols, rows);
    else
        LAUNCH_N("k_blur_box", (k_blur<O3R_BLUR_BOX>), g, kBlurRows, sm, ctx->d_blurjobs.as<BlurJob>(), rows, cols, kernel, 0, 0, cols, rows);
    CU(cudaMemcpy2DAsync(dst, dst_step, ctx->d_blur.p, step, cols, rows, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return O3R_OK;
}

// Pre-reduces the last batch on the combined grid and buckets the partial cells by owner.  Nothing here waits for the GPU:
// launches are sized by the host's bound of the cycle's cell count, the exact count stays on the device.
// info (device, world + 8 words, may be null): counts per owner, the cycle's cell range, the cell count.
static int exchange_pack_impl(o3r_ctx* ctx, int world, o3r_cell* send_dev, size_t cap, uint32_t* info_dev, uint32_t** ocnt_out) {
    if (ctx->retain()) return ctx->fail(O3R_ERR_UNSUPPORTED, "exchange needs O3R_MERGE_ACCUMULATE");
    uint32_t* cnt = ctx->counters.as<uint32_t>();
    CU(ctx->okeys.ensure((size_t)kMaxPasses * kRsBins * 4 + sizeof(SortPlan)));
    uint32_t* ocnt = ctx->okeys.as<uint32_t>();   // laid out like ghist: the owner histogram is pass 0's
    if (ocnt_out) *ocnt_out = ocnt;
    ZERO(ocnt, (size_t)kMaxPasses * kRsBins * 4);
    const int none[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
    const int* bbp = ctx->last_has_cellbb ? ctx->last_cellbb : nullptr;
    if (!ctx->last_is_vox || ctx->last_total == 0) {   // nothing to send: header only
        ctx->n_cyc_ub = 0;
        ZERO(cnt + CNT_CYC, 4);
        if (info_dev)
            LAUNCH(k_pack_cells, 1, kThreads, 0, cnt + CNT_CYC, (const uint32_t*)nullptr, (const uint64_t*)nullptr, (const float4*)nullptr,
                   (const uint4*)nullptr, send_dev, ocnt, (uint32_t)world, none[0], none[1], none[2], none[3], none[4], none[5], info_dev);
        return O3R_OK;
    }
    int rc;
    if (ctx->last_has_partials) {
        AccItemsCells items{ctx->partials.as<o3r_cell>(), nullptr, nullptr};
        rc = acc_build_cycle(ctx, items, ctx->last_partials, false, bbp);
    } else {
        AccItemsPts items{ctx->vox.as<float4>(), nullptr, nullptr};
        rc = acc_build_cycle(ctx, items, ctx->last_total, false, bbp);
    }
    if (rc) return rc;
    const size_t nub = ctx->n_cyc_ub;
    if (cap < nub) return ctx->fail(O3R_ERR_CAPACITY, "send buffer smaller than o3r_exchange_bound()");
    // one stable radix pass on the owner id buckets the (key-sorted) cells by destination rank
    SortU32 sb;
    rc = carve_sort_u32(ctx, nub, sb);
    if (rc) return rc;
    SortPlan* plan = reinterpret_cast<SortPlan*>(ocnt + kMaxPasses * kRsBins);
    SortPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.active_mask = 1; pl.final_parity = 1; pl.n_active = 1; pl.n_passes = 1; pl.bits[0] = 8;
    { int rcu = upload_small(ctx, plan, &pl, sizeof(pl)); if (rcu) return rcu; }
    CU(ctx->seg2.ensure(16));
    const uint32_t g = std::min<uint32_t>(std::max(1u, cdiv(nub, kThreads)), 148 * 8);
    LAUNCH(k_owner, g, kThreads, 0, cnt + CNT_CYC, ctx->ckey.as<uint64_t>(), (uint32_t)world, sb.k0, sb.v0, ocnt, ctx->seg2.as<uint32_t>());
    rc = sort_pairs<uint32_t>(ctx, sb.k0, sb.k1, sb.v0, sb.v1, ctx->seg2.as<uint32_t>(), 1, std::max<size_t>(nub, 1), plan, 1, 0, ocnt,
                              0 /* raw owner counts: they are also the exchange header */);
    if (rc) return rc;
    const int* bb = bbp ? bbp : none;
    LAUNCH(k_pack_cells, g, kThreads, 0, cnt + CNT_CYC, sb.v1, ctx->ckey.as<uint64_t>(), ctx->cacc.as<float4>(), ctx->crgb.as<uint4>(),
           send_dev, ocnt, (uint32_t)world, bb[0], bb[1], bb[2], bb[3], bb[4], bb[5], info_dev);
    return O3R_OK;
}

size_t o3r_exchange_bound(o3r_ctx* ctx) {
    if (!ctx) return 0;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->last_is_vox || ctx->last_total == 0) return 0;
    // cells the next pack can emit: no more than the batch has records, no more than its cell range holds
    size_t n = ctx->last_has_partials ? ctx->last_partials : ctx->last_total;
    if (ctx->last_has_cellbb) {
        const int* b = ctx->last_cellbb;
        const double cells = ((double)b[3] - b[0] + 1) * ((double)b[4] - b[1] + 1) * ((double)b[5] - b[2] + 1);
        if (cells > 0 && cells < (double)n) n = (size_t)cells;
    }
    return n;
}

int o3r_exchange_pack_dev(o3r_ctx* ctx, int world, o3r_cell* send_dev, size_t cap, uint32_t* info_dev) {
    if (!ctx || world < 1 || world > 256 || !info_dev) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    return exchange_pack_impl(ctx, world, send_dev, cap, info_dev, nullptr);
}

int o3r_exchange_pack(o3r_ctx* ctx, int world, o3r_cell* send_dev, size_t cap, uint32_t* counts) {
    if (!ctx || world < 1 || world > 256 || !counts) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    uint32_t* ocnt = nullptr;
    int rc = exchange_pack_impl(ctx, world, send_dev, cap, nullptr, &ocnt);
    if (rc) return rc;
    CU(cudaMemcpyAsync(counts, ocnt, (size_t)world * 4, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(ctx->h_counters + CNT_CYC, ctx->counters.as<uint32_t>() + CNT_CYC, 4, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    ctx->n_cyc = ctx->h_counters[CNT_CYC];
    return O3R_OK;
}

static int exchange_merge_impl(o3r_ctx* ctx, const o3r_cell* recv_dev, size_t n, const int* bb) {
    if (ctx->retain()) return ctx->fail(O3R_ERR_UNSUPPORTED, "exchange needs O3R_MERGE_ACCUMULATE");
    if (n == 0) return O3R_OK;
    AccItemsCells items{recv_dev, nullptr, nullptr};
    int rc = acc_build_cycle(ctx, items, n, true, bb);
    if (rc) return rc;
    return acc_apply_cycle(ctx);
}

int o3r_exchange_merge(o3r_ctx* ctx, const o3r_cell* recv_dev, size_t n) {
    if (!ctx || (n && !recv_dev)) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    return exchange_merge_impl(ctx, recv_dev, n, nullptr);
}

int o3r_exchange_merge_bb(o3r_ctx* ctx, const o3r_cell* recv_dev, size_t n, const int bb[6]) {
    if (!ctx || (n && !recv_dev) || !bb) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    return exchange_merge_impl(ctx, recv_dev, n, bb);
}


// ---- pre-pass ---------------------------------------------------------------------------------------------
namespace {

// symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations: AtA = V diag(w) V^T (what cv::invert(AtA, DECOMP_SVD)
// needs for a symmetric positive semi-definite matrix)
void sym3_eigen(double A[3][3], double V[3][3], double w[3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
    for (int sweep = 0; sweep < 64; ++sweep) {
        const double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
        if (off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - sn * akq;
                    A[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - sn * aqk;
                    A[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq;
                    V[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

// getMean + getVariance (pose_functions.cpp:987-1028) of the ROI; the sample source is already on the device
int roi_variance(o3r_ctx* ctx, const uint8_t* d_disp, size_t dstep, const uint8_t* d_labels, size_t lstep,
                 const double* d_coef, int n_planes, double* variance) {
    const o3r_params& p = ctx->p;
    const int nrow = p.rows - 2 * p.bounding_box, ncol = p.cols - p.bounding_box - p.cols_start_aft_cutout;
    if (nrow <= 0 || ncol <= 0) return ctx->fail(O3R_ERR_INVALID, "empty ROI");
    CU(ctx->pp_rows.ensure((size_t)nrow * 8));
    std::vector<double> rows_h(nrow);
    double mean = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        LAUNCH(k_row_moment, nrow, kThreads, 0, d_disp, dstep, d_labels, lstep, d_coef, n_planes, p.cols,
               p.cols_start_aft_cutout, p.bounding_box, p.min_disparity, pass, mean, ctx->pp_rows.as<double>());
        CU(cudaMemcpyAsync(rows_h.data(), ctx->pp_rows.p, (size_t)nrow * 8, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        double s = 0.0;
        for (int r = 0; r < nrow; ++r) s += rows_h[r];
        if (pass == 0) mean = s / ((double)nrow * ncol);            // :1004 (divides by ALL ROI pixels)
        else *variance = s / ((double)nrow * ncol - 1);             // :1026
    }
    return O3R_OK;
}

int stage_plane_u8(o3r_ctx* ctx, DevBuf& buf, const uint8_t* src, size_t src_step, size_t* step_out) {
    const o3r_params& p = ctx->p;
    const size_t step = ((size_t)p.cols + 15) & ~(size_t)15;
    CU(buf.ensure(step * p.rows));
    CU(cudaMemcpy2DAsync(buf.p, step, src, src_step, p.cols, p.rows, cudaMemcpyHostToDevice, ctx->st));
    *step_out = step;
    return O3R_OK;
}

}  // namespace

int o3r_disp_variance(o3r_ctx* ctx, const uint8_t* disp, size_t disp_step, double* variance) {
    if (!ctx || !disp || !variance) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    size_t step;
    int rc = stage_plane_u8(ctx, ctx->pp_disp, disp, disp_step, &step);
    if (rc) return rc;
    return roi_variance(ctx, ctx->pp_disp.as<uint8_t>(), step, nullptr, 0, nullptr, 0, variance);
}

int o3r_plane_fit(o3r_ctx* ctx, const uint8_t* labels, size_t labels_step, const uint8_t* disp, size_t disp_step,
                  double* coef, int coef_cap, int* n_planes, double* variance) {
    if (!ctx || !labels || !disp || !coef || !n_planes) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    const o3r_params& p = ctx->p;
    size_t lstep, dstep;
    int rc = stage_plane_u8(ctx, ctx->pp_labels, labels, labels_step, &lstep);
    if (rc) return rc;
    rc = stage_plane_u8(ctx, ctx->pp_disp, disp, disp_step, &dstep);
    if (rc) return rc;
    const size_t sum_bytes = (size_t)256 * kLabSums * 8;
    CU(ctx->pp_sums.ensure(sum_bytes));
    ZERO(ctx->pp_sums.p, sum_bytes);
    LAUNCH(k_label_sums, cdiv(p.rows, 8), kThreads, 0, ctx->pp_labels.as<uint8_t>(), lstep, ctx->pp_disp.as<uint8_t>(), dstep,
           p.rows, p.cols, p.cols_start_aft_cutout, p.bounding_box, ctx->pp_sums.as<unsigned long long>());
    std::vector<unsigned long long> S((size_t)256 * kLabSums);
    CU(cudaMemcpyAsync(S.data(), ctx->pp_sums.p, sum_bytes, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    int np = 0;
    for (int cluster = 1; cluster < 256; ++cluster) {   // :907 (labels are 8-bit); stops at the first absent label (:923)
        const unsigned long long* s = &S[(size_t)cluster * kLabSums];
        if (s[0] == 0) break;
        if (cluster > coef_cap) return ctx->fail(O3R_ERR_CAPACITY, "coef buffer too small");
        double* c = coef + 3 * (cluster - 1);
        c[0] = c[1] = c[2] = 0.0;
        np = cluster;
        if (s[1] == 0) continue;   // :939-940: no pixel inside the ROI, the label keeps 0.0
        // :957-965: x = inv_SVD(AtA) * At * b with AtA = [[Sxx Sxy Sx][Sxy Syy Sy][Sx Sy n]], At*b = [Sxd Syd Sd]
        double A[3][3] = {{(double)s[4], (double)s[5], (double)s[2]}, {(double)s[5], (double)s[6], (double)s[3]},
                          {(double)s[2], (double)s[3], (double)s[1]}};
        double V[3][3], w[3];
        sym3_eigen(A, V, w);
        const double thr = DBL_EPSILON * 2 * (std::fabs(w[0]) + std::fabs(w[1]) + std::fabs(w[2]));
        const double rhs[3] = {(double)s[7], (double)s[8], (double)s[9]};
        for (int i = 0; i < 3; ++i) {
            double acc = 0;
            for (int j = 0; j < 3; ++j) {
                double inv_ij = 0;
                for (int e = 0; e < 3; ++e)
                    if (std::fabs(w[e]) > thr) inv_ij += V[i][e] * V[j][e] / w[e];
                acc += inv_ij * rhs[j];
            }
            c[i] = acc;
        }
    }
    *n_planes = np;
    if (variance) {   // getVariance(new_disp_img, true), :973; the caller applies the > 3 gate (:975-982)
        CU(ctx->pp_coef.ensure(std::max(np, 1) * 24));
        if (np) { int rcu = upload_small(ctx, ctx->pp_coef.p, coef, (size_t)np * 24); if (rcu) return rcu; }
        return roi_variance(ctx, nullptr, 0, ctx->pp_labels.as<uint8_t>(), lstep, ctx->pp_coef.as<double>(), np, variance);
    }
    return O3R_OK;
}


int o3r_sor(o3r_ctx* ctx, const o3r_point* pts, size_t n, int mean_k, double stddev_mul, uint8_t* keep, float* dist,
            double* threshold) {
    if (!ctx || (n && !pts) || !keep) return O3R_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->p.device));
    if (n == 0) return O3R_OK;
    if (n >= (1ull << 32)) return ctx->fail(O3R_ERR_INVALID, "cloud too large for 32-bit indices");
    CU(ctx->pts.ensure(n * 16));
    CU(ctx->seg2.ensure(16));
    ctx->last_n = 0; ctx->last_total = 0; ctx->last_has_partials = false;
    CU(cudaMemcpyAsync(ctx->pts.p, pts, n * 16, cudaMemcpyHostToDevice, ctx->st));
    const uint32_t seg_h[2] = {0u, (uint32_t)n};
    { int rcu = upload_small(ctx, ctx->seg2.p, seg_h, 8); if (rcu) return rcu; }
    SortU32 sb;
    int rc = carve_sort_u32(ctx, n, sb);
    if (rc) return rc;
    rc = sor_filter(ctx, sb, ctx->pts.as<float4>(), ctx->seg2.as<uint32_t>(), 1, n, mean_k, stddev_mul);
    if (rc) return rc;
    std::vector<float> d(n);
    double thr = 0;
    CU(cudaMemcpyAsync(d.data(), ctx->sor_dist.p, n * 4, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(&thr, ctx->sor_thr.p, 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    for (size_t i = 0; i < n; ++i) keep[i] = !((double)d[i] > thr);
    if (dist) memcpy(dist, d.data(), n * 4);
    if (threshold) *threshold = thr;
    return O3R_OK;
}

}  // extern "C"
