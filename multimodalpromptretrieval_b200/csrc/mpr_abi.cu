// C-ABI entry points (include/mpr_b200.h): argument validation, launch planning, TMA descriptor encoding and
// kernel launches.  No device synchronisation (except where the caller asks for it in mpr_retrieve_host), no persistent
// device allocations beyond one error word.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/mpr_b200.h"
#include "bank_build.cuh"
#include "embed_gather.cuh"
#include "merge_topk.cuh"
#include "prompt_gather.cuh"
#include "scan_topk.cuh"
#include "tail.cuh"

using namespace mpr;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct mpr_context {
    int device = -1;
    int num_sms = 0;
    int* d_err = nullptr;
    unsigned long long* d_dbg = nullptr;    // debug counters + timeline, allocated only with MPR_DEBUG_COUNTERS=1|2
    int dbg_counters = 0;                   // 1: the event counters are live (their atomics distort timings); 2: timeline only
    PFN_encodeTiled encode = nullptr;
    std::vector<cudaEvent_t> prof_events;   // start/stop pairs, used only between mpr_profile_begin/end
    int cand_cap_override = 0;              // MPR_CAND_CAP=10..16 forces the pending-buffer depth
    int epi_groups = 0;                     // MPR_EPI_GROUPS=1|2 forces the epilogue group count (0 = planner's choice)
    int stage_subs = 4;                     // max 64-wide K sub-chunks per ring stage (MPR_STAGE_SUBS=1|2|4)
    int use_q_tmem = 1;                     // q-tile as TMEM A operand when D <= 512 (MPR_NO_QTMEM=1 disables)
    int use_cluster = 1;                    // CTA-pair TMA multicast in the tensor-bound regime (MPR_NO_CLUSTER=1 disables)
    int dynamic_tiles = 1;                  // tiles pulled from an atomic counter (MPR_STATIC_TILES=1: contiguous ranges)
    int shared_thr = 1;                     // shared admission thresholds (MPR_NO_GTHR=1 disables)
    int fused_tail = 1;                     // single-wave grids finish the step in the scan launch (MPR_NO_FUSED_TAIL=1)
    int cooperative = 1;                    // fused-tail launches are cooperative (MPR_NO_COOP=1: plain launch)
    int use_reg_list = 1;                   // k + skip <= 8: lists in registers (MPR_NO_REGLIST=1: shared memory)
    int first_wait_ns = 16000;              // first tile: bounded wait for the shared thresholds (MPR_FIRST_WAIT_NS; -1 = legacy start)
    int thr_rep_log2 = 2;                   // up to 4 replica words per threshold slot (MPR_THR_REPLICAS=1|2|4)
    int hybrid_min_b = 16;                  // ... for batches beyond this many queries (MPR_HYBRID_MIN_B)
    int use_hybrid = 1;                     // hybrid TMEM + shared-memory q-tile for 512 < D <= 1024 (MPR_NO_HYBRID=1 disables)
    int xchg_mode = 0;                      // MPR_XCHG_MODE tuning bits (tail.cuh XchgParams::mode)
    int pdl = 0;                            // programmatic dependent launch of the scan kernel (MPR_PDL=1)
    int q_coop = 1;                         // warp-cooperative coalesced q-tile fill (MPR_NO_QCOOP=1: a thread per row)
    int tail_floor = 1;                     // pool merge drops keys below the final shared threshold (MPR_NO_TAIL_FLOOR=1)
    unsigned long long xchg_timeout_ns = 60ull * 1000000000ull;
    // workspaces whose control words are known to be zero over `second` leading bytes (the library zeroed them when it
    // first saw the pointer, every launch leaves them zero); most recently used last, at most 16
    std::vector<std::pair<const void*, size_t>> clean_ws;
    cudaEvent_t io_events[4] = {nullptr, nullptr, nullptr, nullptr};   // mpr_retrieve_host with copy streams
    // deferred finish of sharded steps (mpr_retrieve_args.defer_finish): a stream of the library's own, per step j an
    // event "scan j done" and an event "finish j done" (both indexed j & 3), and the epoch words the scan leaves behind
    cudaStream_t fin_stream = nullptr;
    cudaEvent_t fin_scan_done[4] = {nullptr, nullptr, nullptr, nullptr}, fin_done[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t* d_finish_epoch = nullptr;
    unsigned fin_seq = 0;                   // deferred steps queued so far
    bool last_deferred = false;             // the last run_step queued a deferred finish
    int io_turn = 0;
    unsigned launch_seq = 0;                // scan launches so far (debug ring index)
    int last_launches = 0;                  // kernel launches of the last mpr_retrieve
    int prof_used = -1;                     // -1 = profiling off
    int prof_last_n = 0;                    // launches recorded by the last begin/end pair
    char err[512] = {0};
};

constexpr int kDbgRingOff = 8 + 24 * 2048;  // event counters + a 24-slot timeline for up to 2048 CTAs, then
constexpr int kDbgWords = kDbgRingOff + 128; // a ring of [first entry, last exit] for the last 64 launches

static thread_local char g_err[512] = "";

static int fail(mpr_context* h, int code, const char* fmt, ...) {
    char* dst = h ? h->err : g_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(h, call)                                                                              \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(h, MPR_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }
static inline size_t round16(size_t v) { return (v + 15u) & ~size_t(15); }

// Every ABI call runs on the handle's device whatever the caller's current device is, and leaves the caller's current
// device untouched.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

// ------------------------------------------------------------------------------------------------ planning
struct ScanPlan {
    int n_chunks, q_tile, q_box_rows, n_qtiles, n_splits, n_tiles, n_stages, kk_pad, cand_cap, sub_per_stage, n_epi_groups;
    bool q_tmem;       // q-tile in tensor memory (TMEM A operand) instead of shared memory
    bool hybrid;       // 512 < D <= 1024: K-chunks 0..7 of the q-tile in tensor memory, the rest in shared memory
    int n_q_smem;      // K-chunks of the q-tile held in shared memory
    bool reg_list;     // k + skip <= 8: per-query lists in registers (no list / pending memory in shared memory)
    uint32_t smem_bytes;
};

static int make_plan(mpr_context* h, int b, int64_t n_local, int d, int kk, ScanPlan* pl, bool allow_hybrid = true) {
    if (b < 1) return fail(h, MPR_EINVAL, "b must be >= 1 (got %d)", b);
    if (n_local < 0 || n_local >= (1ll << 31) - kTileRows)      // 0 = a rank whose shard is empty still takes part in the exchange
        return fail(h, MPR_EINVAL, "n_local out of range (got %lld)", static_cast<long long>(n_local));
    if (d < 64 || d > 4096 || d % 64 != 0) return fail(h, MPR_EINVAL, "d must be a multiple of 64 in [64, 4096] (got %d)", d);
    if (kk < 1 || kk > MPR_MAX_KK) return fail(h, MPR_EINVAL, "k + skip must be in [1, %d] (got %d)", MPR_MAX_KK, kk);

    pl->n_chunks = d / kChunkK;
    pl->reg_list = h->use_reg_list && kk <= 8;
    pl->kk_pad = pl->reg_list ? 0 : pow2_ceil(kk);
    pl->n_tiles = static_cast<int>((n_local + kTileRows - 1) / kTileRows);
    // D <= 512: the q-tile (128 x D bf16) fits 256 TMEM columns next to two 128-column accumulators
    pl->q_tmem = h->use_q_tmem && d <= 512;
    // 512 < D <= 1024: a q-tile in shared memory alone would have to shrink to 64 queries (M = 64 MMAs at half the
    // tensor rate, the bank read once more per extra q-tile) and feeds both MMA operands through the shared-memory port.
    // The hybrid q-tile keeps up to 128 queries resident: the first 512 dims in tensor memory, the remaining <= 512 in
    // shared memory (<= 128 KiB).  Measured on 1 M x 1024: faster from 17 queries on (B = 64: 400 -> 320 us).
    pl->hybrid = allow_hybrid && h->use_hybrid && h->use_q_tmem && h->q_coop && d > 512 && d <= 1024 && b > h->hybrid_min_b;
    if (pl->hybrid) pl->q_tmem = true;
    pl->n_q_smem = pl->hybrid ? pl->n_chunks - kQTmemChunks : (pl->q_tmem ? 0 : pl->n_chunks);
    int q_tile_max = 128;
    while (!pl->q_tmem && q_tile_max > 8 && static_cast<long long>(q_tile_max) * d * 2 > 131072) q_tile_max >>= 1;
    // Shared memory is split between the resident q-tile, the per-query lists and the bank ring.  Reading the bank
    // twice (two q-tiles) costs far more than a shallower ring, so the q-tile is only halved when fewer than 3 stages
    // (48 KiB in flight per SM) would remain even with the smallest pending buffers.
    int stages = 0;
    for (;;) {
        if (b <= q_tile_max) {
            pl->q_tile = b;
            pl->n_qtiles = 1;
            pl->q_box_rows = pow2_ceil(b) < 8 ? 8 : pow2_ceil(b);
        } else {
            pl->q_tile = q_tile_max;
            pl->n_qtiles = (b + q_tile_max - 1) / q_tile_max;
            pl->q_box_rows = q_tile_max;
        }
        // D <= 512: nothing of Q in shared memory; hybrid: 64- or 128-row slabs for the dims beyond 512
        if (pl->q_tmem) {
            pl->q_box_rows = 0;
            if (pl->hybrid) {      // slab rows: the batch rounded up to 16, and >= 32 KiB in all (the fill's scratch)
                int rows = (std::min(b, 128) + 15) / 16 * 16;
                const int min_rows = ((256 + pl->n_q_smem - 1) / pl->n_q_smem + 15) / 16 * 16;
                pl->q_box_rows = std::min(128, std::max(rows, min_rows));
            }
        }
        // Two epilogue groups (two lists per query) unless the second group's lists would starve the bank ring
        // (SS mode keeps 64-128 KiB of Q in shared memory, and a k+s = 32 list is 392 B per query).  With the shared
        // admission thresholds list maintenance is off the critical path for every k, so k no longer decides this.
        // (deeper pending buffers were measured to HURT: 32 slots -> +25 % at k+s = 16/32, because the admission
        // threshold only moves at a flush and a stale threshold admits many more candidates)
        const int caps1[] = {16, 14, 12, 10, 10}, caps2[] = {16, 14, 12, 10};
        // register lists: only the pending buffer lives in shared memory; beside a hybrid q-tile (128 KiB of shared
        // memory) the smallest buffer buys a fourth ring stage, which the L2-latency-bound B-operand stream needs more
        const int caps0[] = {pl->hybrid ? 10 : 16};
        auto units_for = [&](int cap, int groups) {   // 16 KiB ring units left beside the resident q-tile and the lists
            const ScanSmemLayout fixed = scan_smem_layout(pl->n_q_smem, pl->q_box_rows, pl->kk_pad, cap, 0, 1, groups);
            return (kMaxSmem - 1024 - static_cast<int>(fixed.total)) / kStageBytes;
        };
        pl->n_epi_groups = (!pl->reg_list && units_for(16, 2) < 6) ? 1 : 2;
        // beside a full-height hybrid q-tile a fifth ring stage is worth more than the second epilogue group: a D = 1024
        // tile streams for ~5 us, one group needs ~1.5 us of it
        if (pl->hybrid && pl->q_box_rows > 64) pl->n_epi_groups = 1;
        if (h->epi_groups == 1 || (h->epi_groups == 2 && units_for(10, 2) >= 3)) pl->n_epi_groups = h->epi_groups;
        const int* caps = pl->reg_list ? caps0 : pl->n_epi_groups == 1 ? caps1 : caps2;
        const int n_caps = pl->reg_list ? 1 : pl->n_epi_groups == 1 ? 5 : 4;
        pl->cand_cap = caps[n_caps - 1];
        stages = units_for(pl->cand_cap, pl->n_epi_groups);
        for (int want : {8, 6, 3}) {     // 128 KiB in flight per SM if any pending depth allows it, else 96, else 48
            bool found = false;
            for (int c = 0; c < n_caps && !found; ++c)
                if (units_for(caps[c], pl->n_epi_groups) >= want) {
                    pl->cand_cap = caps[c];
                    stages = units_for(caps[c], pl->n_epi_groups);
                    found = true;
                }
            if (found) break;
        }
        if (!pl->reg_list && h->cand_cap_override >= 10 && h->cand_cap_override <= kCandCapMax && units_for(h->cand_cap_override, pl->n_epi_groups) >= 3) {
            pl->cand_cap = h->cand_cap_override;      // tuning knob (MPR_CAND_CAP)
            stages = units_for(pl->cand_cap, pl->n_epi_groups);
        }
        if (stages >= 3 || q_tile_max <= 32 || pl->q_tmem || pl->q_box_rows < q_tile_max) break;
        q_tile_max >>= 1;
    }
    if (pl->hybrid) {
        // the ring needs at least three 16 KiB stages (the q-tile fill's scratch is the shared-memory half of the q-tile
        // itself: n_q_smem >= 2 slabs of 16 KiB)
        if (stages < 3 || pl->n_q_smem < 2) return make_plan(h, b, n_local, d, kk, pl, false);
    }
    if (stages < 2) return fail(h, MPR_EINVAL, "shape does not fit shared memory (d=%d, kk=%d)", d, kk);
    // Group 64-wide K sub-chunks into 32 / 64 KiB ring stages (one barrier round-trip per 8 / 16 MMAs — the MMA warp is
    // otherwise bound by its own barrier + issue overhead) as long as at least three stages remain.
    pl->sub_per_stage = 1;
    if (h->stage_subs >= 2 && pl->n_chunks >= 2 && stages >= 6) pl->sub_per_stage = 2;
    if (h->stage_subs >= 4 && pl->n_chunks >= 4 && stages >= 12) pl->sub_per_stage = 4;
    stages /= pl->sub_per_stage;
    if (stages > kMaxStages) stages = kMaxStages;
    pl->n_stages = stages;
    // items = n_splits * n_qtiles should be a whole number of waves over the SMs
    const int g = std::gcd(h->num_sms, pl->n_qtiles);
    pl->n_splits = h->num_sms / g;
    if (pl->n_splits > pl->n_tiles) pl->n_splits = pl->n_tiles;
    if (pl->n_splits < 1) pl->n_splits = 1;
    pl->smem_bytes = scan_smem_layout(pl->n_q_smem, pl->q_box_rows, pl->kk_pad, pl->cand_cap, stages, pl->sub_per_stage, pl->n_epi_groups).total + 1024u;
    return MPR_OK;
}

// Workspace = [control block | tile counters | shared thresholds] (all-zero between launches) + partial lists.
struct WsLayout {
    size_t tile_ctr_off, gthr_off, zero_bytes, part_off, total;
    int ns, rep_log2;
};

static WsLayout ws_layout(const mpr_context* h, const ScanPlan& pl, int b, int kk) {
    WsLayout w;
    w.ns = (kk + 3) & ~3;
    w.rep_log2 = 0;
    const int max_words = pl.reg_list ? 16 : 32;       // what the scan variant holds in registers (kThrN)
    while (w.rep_log2 < h->thr_rep_log2 && (w.ns << (w.rep_log2 + 1)) <= max_words) ++w.rep_log2;
    w.tile_ctr_off = 16;
    w.gthr_off = round16(w.tile_ctr_off + sizeof(uint32_t) * static_cast<size_t>(pl.n_qtiles));
    w.zero_bytes = w.gthr_off + sizeof(uint32_t) * static_cast<size_t>(b) * (w.ns << w.rep_log2);
    w.part_off = round16(w.zero_bytes);
    w.total = w.part_off + static_cast<size_t>(pl.n_splits) * kEpiGroups * b * kk * sizeof(uint64_t);
    return w;
}

static int encode_2d(mpr_context* h, CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols,
                     uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {kChunkK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = h->encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(h, MPR_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box_rows=%u", static_cast<int>(r),
                    static_cast<unsigned long long>(rows), static_cast<unsigned long long>(cols), box_rows);
    return MPR_OK;
}

template <typename Kern>
static cudaError_t launch_kernel(Kern kern, dim3 grid, uint32_t smem, cudaStream_t st, int cluster, bool cooperative,
                                 bool pdl, const CUtensorMap& tq, const CUtensorMap& tb, const ScanParams& p,
                                 const TailParams& t) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kScanThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[3];
    int n = 0;
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (cooperative) {
        attr[n].id = cudaLaunchAttributeCooperative;
        attr[n].val.cooperative = 1;
        ++n;
    }
    if (pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, tq, tb, p, t);
}

// One retrieval step on the device: [kernel 1] -> scan (+ fused tail) -> [stand-alone tail].
template <bool kDump>
static int run_step(mpr_context* h, const mpr_retrieve_args& a, float* dump, cudaStream_t st) {
    const int b = a.b, kk = a.kk, d = a.d;
    ScanPlan pl;
    const int saved_reg = h->use_reg_list;
    if (kDump) h->use_reg_list = 0;            // the dump epilogue exists for the shared-memory-list variant only
    // the hybrid q-tile takes prepared bf16 queries or raw ones that need no normalisation (the warps fill both halves);
    // normalised raw queries go through kernel 1 into the caller's scratch first
    const bool can_prepare = a.q0 == nullptr || a.q_scratch != nullptr || !a.normalise;
    int rc = make_plan(h, b, a.n_local, d, kk, &pl, !kDump && can_prepare);
    h->use_reg_list = saved_reg;
    if (rc) return rc;
    const WsLayout wl = ws_layout(h, pl, b, kk);
    if (a.workspace_bytes < wl.total)
        return fail(h, MPR_EWORKSPACE, "workspace too small: %zu < %zu", a.workspace_bytes, wl.total);
    h->last_launches = 0;

    // ---- which kernel variant
    bool raw = a.q0 != nullptr;
    const uint16_t* q_bf16 = a.q_bf16;
    // tensor-bound regime with an even number of q-tiles: CTA pairs share each bank chunk by TMA multicast
    bool pair = !kDump && !pl.q_tmem && h->use_cluster && pl.n_qtiles >= 2 && pl.n_qtiles % 2 == 0;
    if (raw && (pair || (pl.hybrid && a.normalise))) {
        if (a.q_scratch) {     // these variants take prepared queries: one kernel-1 launch into the caller's scratch
            const int threads = 256, rows_per_block = threads / 32;
            bank_build_kernel<<<(b + rows_per_block - 1) / rows_per_block, threads, 0, st>>>(
                a.q0, a.d0, a.q1, a.q1 ? a.d1 : 0, a.q_dtype, b, a.normalise, a.q_scratch, a.out_q_bias);
            CUDA_TRY(h, cudaGetLastError());
            ++h->last_launches;
            q_bf16 = a.q_scratch;
            raw = false;
        } else {
            pair = false;
        }
    }
    if (raw && !pl.q_tmem && d > 2048)
        return fail(h, MPR_EINVAL, "raw queries are supported up to D = 2048 (got %d); prepare them with mpr_bank_build", d);
    if (!raw && !q_bf16) return fail(h, MPR_EINVAL, "no queries: q0 and q_bf16 are both null");

    CUtensorMap tq, tb;
    if ((pl.q_tmem && !pl.hybrid) || raw) memset(&tq, 0, sizeof(tq));       // the warps bring the q-tile in; no Q tensor map
    else rc = encode_2d(h, &tq, q_bf16, static_cast<uint64_t>(b), static_cast<uint64_t>(d), pl.q_box_rows);
    if (rc) return rc;
    if (a.n_local > 0) rc = encode_2d(h, &tb, a.bank, static_cast<uint64_t>(a.n_local), static_cast<uint64_t>(d), pair ? kTileRows / 2 : kTileRows);
    else memset(&tb, 0, sizeof(tb));                         // no tiles: the bank map is never dereferenced
    if (rc) return rc;

    const dim3 grid(pl.n_splits * pl.n_qtiles);
    const bool fused_tail = !kDump && !pair && h->fused_tail && static_cast<int>(grid.x) <= h->num_sms;
    unsigned char* ws = static_cast<unsigned char*>(a.workspace);
    if (!kDump) {
        // control words must be zero before the first launch on this workspace; every launch leaves them zero again
        size_t known = 0;
        for (size_t i = 0; i < h->clean_ws.size(); ++i)
            if (h->clean_ws[i].first == ws) {
                known = h->clean_ws[i].second;
                h->clean_ws.erase(h->clean_ws.begin() + static_cast<long>(i));
                break;
            }
        if (wl.zero_bytes > known) CUDA_TRY(h, cudaMemsetAsync(ws, 0, wl.zero_bytes, st));
        if (h->clean_ws.size() >= 16) h->clean_ws.erase(h->clean_ws.begin());
        // after this launch exactly the current shape's control region is guaranteed zero (a larger region of an earlier
        // shape may since have been used for partial lists)
        h->clean_ws.emplace_back(ws, wl.zero_bytes);
    }

    ScanParams p;
    p.b_total = b;
    p.n_local = static_cast<int>(a.n_local);
    p.n_chunks = pl.n_chunks;
    p.n_q_smem = pl.n_q_smem;
    p.kk = kk;
    p.kk_pad = pl.kk_pad;
    p.cand_cap = pl.cand_cap;
    p.q_tile = pl.q_tile;
    p.q_box_rows = pl.q_box_rows;
    p.n_qtiles = pl.n_qtiles;
    p.n_splits = pl.n_splits;
    p.n_tiles = pl.n_tiles;
    p.n_stages = pl.n_stages;
    p.sub_per_stage = pl.sub_per_stage;
    p.n_epi_groups = pl.n_epi_groups;
    p.idx_base = static_cast<uint32_t>(a.idx_base);
    p.bank_policy = pl.n_qtiles == 1 ? ptx::kEvictFirst : ptx::kEvictNormal;
    p.bias = a.bias;
    p.q = q_bf16;
    p.d = d;
    p.qsrc0 = raw ? a.q0 : nullptr;
    p.qsrc1 = raw ? a.q1 : nullptr;
    p.qd0 = a.d0;
    p.qd1 = a.q1 ? a.d1 : 0;
    p.q_dtype = a.q_dtype;
    p.q_normalise = a.normalise;
    p.q_bias_out = raw ? a.out_q_bias : nullptr;
    p.part_keys = reinterpret_cast<uint64_t*>(ws + wl.part_off);
    p.tile_ctr = (!kDump && !pair && h->dynamic_tiles) ? reinterpret_cast<uint32_t*>(ws + wl.tile_ctr_off) : nullptr;
    p.gthr = (!kDump && h->shared_thr) ? reinterpret_cast<uint32_t*>(ws + wl.gthr_off) : nullptr;
    p.ns = wl.ns;
    p.thr_rep_log2 = wl.rep_log2;
    p.fused_tail = fused_tail ? 1 : 0;
    {
        const ScanSmemLayout lay = scan_smem_layout(pl.n_q_smem, pl.q_box_rows, pl.kk_pad, pl.cand_cap, pl.n_stages,
                                                    pl.sub_per_stage, pl.n_epi_groups);
        p.q_coop = (h->q_coop && pl.q_tmem && !(raw && a.normalise) &&
                    (pl.hybrid || lay.bias_off - lay.list_off >= 8u * 4096u)) ? 1 : 0;
    }
    // waiting for the slots only pays when every slot is fed by some list that gets a tile right away
    p.first_wait_ns = (h->first_wait_ns > 0 && (pl.n_splits < wl.ns || pl.n_tiles < 2 * pl.n_splits * pl.n_epi_groups))
                          ? 0 : h->first_wait_ns;
    p.dbg = h->dbg_counters ? h->d_dbg : nullptr;
    p.dbg_ts = h->d_dbg ? h->d_dbg + 8 : nullptr;
    p.dbg_ring = h->d_dbg ? h->d_dbg + kDbgRingOff : nullptr;
    p.launch_seq = h->launch_seq++;
    p.dump = dump;
    p.err = h->d_err;

    TailParams t;
    memset(&t, 0, sizeof(t));
    t.b = b;
    t.kk = kk;
    t.part_keys = p.part_keys;
    t.n_lists = pl.n_splits * kEpiGroups;
    t.ctrl = reinterpret_cast<uint32_t*>(ws);
    t.tile_ctr = reinterpret_cast<uint32_t*>(ws + wl.tile_ctr_off);
    t.n_tile_ctr = pl.n_qtiles;
    t.gthr = p.gthr;
    t.ns = wl.ns;
    t.thr_rep_log2 = wl.rep_log2;
    t.use_floor = h->tail_floor;
    t.out_keys = a.out_keys;
    t.out_score = a.out_score;
    t.out_idx = a.out_idx;
    t.status = a.status;
    t.xchg.world = a.world > 1 ? a.world : 1;
    t.xchg.rank = a.world > 1 ? a.rank : 0;
    t.xchg.cap = a.xchg_cap;
    // deferred finish: only for sharded single-launch steps
    const bool defer = !kDump && a.world > 1 && a.defer_finish && fused_tail;
    h->last_deferred = false;
    unsigned fin_j = 0;
    if (defer) {
        if (!h->fin_stream) {
            CUDA_TRY(h, cudaStreamCreateWithFlags(&h->fin_stream, cudaStreamNonBlocking));
            for (int i = 0; i < 4; ++i) {
                CUDA_TRY(h, cudaEventCreateWithFlags(&h->fin_scan_done[i], cudaEventDisableTiming));
                CUDA_TRY(h, cudaEventCreateWithFlags(&h->fin_done[i], cudaEventDisableTiming));
            }
            CUDA_TRY(h, cudaMalloc(&h->d_finish_epoch, 4 * sizeof(uint32_t)));
            CUDA_TRY(h, cudaMemset(h->d_finish_epoch, 0, 4 * sizeof(uint32_t)));
        }
        fin_j = h->fin_seq++;
        // step j is queued only after the finish of step j-2 has run (bounds what is in flight; see tail.cuh)
        if (fin_j >= 2) CUDA_TRY(h, cudaStreamWaitEvent(st, h->fin_done[(fin_j - 2) & 3], 0));
        t.defer_xchg = 1;
        t.finish_epoch = h->d_finish_epoch + (fin_j & 3);
    }
    t.xchg.timeout_ns = h->xchg_timeout_ns;
    t.xchg.mode = h->xchg_mode;
    if (a.world > 1)
        for (int r = 0; r < a.world; ++r) t.xchg.peers.buf[r] = static_cast<unsigned char*>(a.peer_bufs[r]);
    if (a.answer_id) {
        PromptParams& pp = t.prompt;
        pp.idx = nullptr; pp.b = b; pp.kk = kk; pp.skip = a.skip;
        pp.answer_id = a.answer_id; pp.bucket_lut = a.bucket_lut;
        pp.prefix_ids = a.prefix_ids; pp.prefix_off = a.prefix_off; pp.seg_ids = a.seg_ids; pp.seg_off = a.seg_off;
        pp.use_quantifier = a.use_quantifier; pp.pad_id = a.pad_id; pp.eos_id = a.eos_id;
        pp.max_len = a.max_len; pp.out_stride = a.out_stride;
        pp.input_ids = reinterpret_cast<long long*>(a.input_ids);
        pp.attention_mask = reinterpret_cast<long long*>(a.attention_mask);
        pp.out_len = a.out_len; pp.maj_answer = a.maj_answer; pp.maj_count = a.maj_count; pp.bucket = a.bucket;
        pp.ret_answer = a.ret_answer;
    }

    const bool prof = h->prof_used >= 0 && 2 * (h->prof_used + 1) <= static_cast<int>(h->prof_events.size());
    if (prof) CUDA_TRY(h, cudaEventRecord(h->prof_events[2 * h->prof_used], st));
    // a plain launch with programmatic stream serialization when asked for (MPR_PDL=1): the grid barrier then relies on
    // the grid being one wave (it is: grid <= SM count, one CTA per SM) instead of the cooperative-launch guarantee
    const bool pdl = h->pdl && !kDump && !pair;
    const bool coop = fused_tail && h->cooperative && !pdl;
    cudaError_t le;
    const uint32_t sm = pl.smem_bytes;
#define MPR_LAUNCH(DUMP, CL, QT, FQ, RL, COOP) \
    launch_kernel(scan_topk_kernel<DUMP, CL, QT, FQ, RL>, grid, sm, st, CL, COOP, pdl, tq, tb, p, t)
    if (pl.hybrid) {
        if (raw && pl.reg_list)  le = launch_kernel(scan_topk_kernel<false, 1, true, true, true, true>, grid, sm, st, 1, coop, pdl, tq, tb, p, t);
        else if (raw)            le = launch_kernel(scan_topk_kernel<false, 1, true, true, false, true>, grid, sm, st, 1, coop, pdl, tq, tb, p, t);
        else if (pl.reg_list)    le = launch_kernel(scan_topk_kernel<false, 1, true, false, true, true>, grid, sm, st, 1, coop, pdl, tq, tb, p, t);
        else                     le = launch_kernel(scan_topk_kernel<false, 1, true, false, false, true>, grid, sm, st, 1, coop, pdl, tq, tb, p, t);
    } else if (pl.reg_list && !kDump) {
        if (pair)                     le = MPR_LAUNCH(false, 2, false, false, true, false);
        else if (pl.q_tmem && raw)    le = MPR_LAUNCH(false, 1, true, true, true, coop);
        else if (pl.q_tmem)           le = MPR_LAUNCH(false, 1, true, false, true, coop);
        else if (raw)                 le = MPR_LAUNCH(false, 1, false, true, true, coop);
        else                          le = MPR_LAUNCH(false, 1, false, false, true, coop);
    } else {
        if (pair)                     le = MPR_LAUNCH(false, 2, false, false, false, false);
        else if (pl.q_tmem && raw)    le = MPR_LAUNCH(false, 1, true, true, false, coop);
        else if (pl.q_tmem)           le = MPR_LAUNCH(kDump, 1, true, false, false, coop);
        else if (raw)                 le = MPR_LAUNCH(false, 1, false, true, false, coop);
        else                          le = MPR_LAUNCH(kDump, 1, false, false, false, coop);
    }
#undef MPR_LAUNCH
    if (le != cudaSuccess) return fail(h, MPR_ECUDA, "scan launch failed: %s", cudaGetErrorString(le));
    ++h->last_launches;
    if (prof) {
        CUDA_TRY(h, cudaEventRecord(h->prof_events[2 * h->prof_used + 1], st));
        ++h->prof_used;
    }
    if (!kDump && !fused_tail) {
        const int blocks = b < h->num_sms * 8 ? b : h->num_sms * 8;      // one block per query, grid-stride beyond that
        tail_kernel<<<blocks, kTailWarps * 32, 0, st>>>(t);
        CUDA_TRY(h, cudaGetLastError());
        ++h->last_launches;
    }
    if (defer) {
        CUDA_TRY(h, cudaEventRecord(h->fin_scan_done[fin_j & 3], st));
        CUDA_TRY(h, cudaStreamWaitEvent(h->fin_stream, h->fin_scan_done[fin_j & 3], 0));
        xchg_finish_kernel<<<(b + kFinishWarps - 1) / kFinishWarps, kFinishWarps * 32, 0, h->fin_stream>>>(t);
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaEventRecord(h->fin_done[fin_j & 3], h->fin_stream));
        ++h->last_launches;
        h->last_deferred = true;
    }
    return MPR_OK;
}

static int validate_retrieve(mpr_context* h, const mpr_retrieve_args* a) {
    if (!a) return fail(h, MPR_EINVAL, "null argument block");
    if (a->b == 0) return MPR_OK;
    if (!a->workspace || (a->n_local > 0 && (!a->bank || !a->bias)))
        return fail(h, MPR_EINVAL, "null pointer (bank, bias or workspace)");
    if (!aligned16(a->bank) || !aligned16(a->workspace) || !aligned16(a->q0) || !aligned16(a->q1) || !aligned16(a->q_bf16) ||
        !aligned16(a->q_scratch))
        return fail(h, MPR_EINVAL, "queries, bank and workspace must be 16-byte aligned");
    if (a->idx_base < 0 || a->idx_base + a->n_local >= 0xFFFFFFFFll) return fail(h, MPR_EINVAL, "global row index exceeds 32 bits");
    if (a->q0) {
        const int d1 = a->q1 ? a->d1 : 0;
        if (a->d0 < 8 || a->d0 % 8 || d1 % 8 || a->d0 + d1 != a->d)
            return fail(h, MPR_EINVAL, "query halves must be multiples of 8 wide and add up to d (d0=%d d1=%d d=%d)", a->d0, d1, a->d);
        if (a->q_dtype < MPR_SRC_F32 || a->q_dtype > MPR_SRC_BF16) return fail(h, MPR_EINVAL, "bad q_dtype %d", a->q_dtype);
    }
    if (a->world > 1) {
        if (a->world > kXchgMaxWorld || a->rank < 0 || a->rank >= a->world)
            return fail(h, MPR_EINVAL, "bad rank/world %d/%d (max world %d)", a->rank, a->world, kXchgMaxWorld);
        if (!a->peer_bufs) return fail(h, MPR_EINVAL, "peer_bufs is null");
        for (int r = 0; r < a->world; ++r)
            if (!a->peer_bufs[r] || !aligned16(a->peer_bufs[r])) return fail(h, MPR_EINVAL, "peer buffer %d is null or unaligned", r);
        if (static_cast<long long>(a->b) * a->kk > a->xchg_cap)
            return fail(h, MPR_EINVAL, "b*kk = %lld exceeds the exchange capacity %d", static_cast<long long>(a->b) * a->kk, a->xchg_cap);
    }
    if (a->answer_id) {
        if (!a->bucket_lut || !a->prefix_ids || !a->prefix_off || !a->seg_ids || !a->seg_off || !a->input_ids ||
            !a->attention_mask || !a->out_len || !a->maj_answer || !a->maj_count || !a->bucket)
            return fail(h, MPR_EINVAL, "prompt stage: null pointer");
        if (a->skip < 0 || a->skip >= a->kk) return fail(h, MPR_EINVAL, "need 0 <= skip < kk (kk=%d skip=%d)", a->kk, a->skip);
        if (a->max_len < 1 || a->out_stride < 1) return fail(h, MPR_EINVAL, "max_len and out_stride must be >= 1");
    }
    return MPR_OK;
}

// ------------------------------------------------------------------------------------------------ exports
extern "C" {

int mpr_abi_version(void) { return MPR_ABI_VERSION; }

const char* mpr_last_error(mpr_handle_t h) { return h ? h->err : g_err; }

int mpr_create(int device, mpr_handle_t* out) {
    if (!out) return fail(nullptr, MPR_EINVAL, "out is null");
    *out = nullptr;
    DeviceGuard guard(device);       // the caller's current device is restored on return
    int cur = -1;
    CUDA_TRY(nullptr, cudaGetDevice(&cur));
    if (cur != device) CUDA_TRY(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, MPR_EARCH, "device %d is sm_%d%d; this library is sm_100a-only (no fallback path)", device,
                    prop.major, prop.minor);
    mpr_context* h = new mpr_context();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        delete h;
        return fail(nullptr, MPR_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    }
    h->encode = reinterpret_cast<PFN_encodeTiled>(fn);
    e = cudaMalloc(&h->d_err, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(h->d_err, 0, sizeof(int));
    auto opt_in = [&](auto kern) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    };
    opt_in(scan_topk_kernel<false, 1, false, false>);
    opt_in(scan_topk_kernel<true, 1, false, false>);
    opt_in(scan_topk_kernel<false, 2, false, false>);
    opt_in(scan_topk_kernel<false, 1, true, false>);
    opt_in(scan_topk_kernel<true, 1, true, false>);
    opt_in(scan_topk_kernel<false, 1, true, true>);
    opt_in(scan_topk_kernel<false, 1, false, true>);
    opt_in(scan_topk_kernel<false, 1, false, false, true>);
    opt_in(scan_topk_kernel<false, 2, false, false, true>);
    opt_in(scan_topk_kernel<false, 1, true, false, true>);
    opt_in(scan_topk_kernel<false, 1, true, true, true>);
    opt_in(scan_topk_kernel<false, 1, false, true, true>);
    opt_in(scan_topk_kernel<false, 1, true, false, true, true>);
    opt_in(scan_topk_kernel<false, 1, true, false, false, true>);
    opt_in(scan_topk_kernel<false, 1, true, true, true, true>);
    opt_in(scan_topk_kernel<false, 1, true, true, false, true>);
    {
        auto flag = [](const char* name) { const char* v = getenv(name); return v && v[0] == '1'; };
        if (flag("MPR_NO_CLUSTER")) h->use_cluster = 0;
        const char* cc = getenv("MPR_CAND_CAP");
        if (cc) h->cand_cap_override = atoi(cc);
        const char* eg = getenv("MPR_EPI_GROUPS");
        if (eg && (eg[0] == '1' || eg[0] == '2')) h->epi_groups = eg[0] - '0';
        const char* ss = getenv("MPR_STAGE_SUBS");
        if (ss && (ss[0] == '1' || ss[0] == '2' || ss[0] == '4')) h->stage_subs = ss[0] - '0';
        if (flag("MPR_NO_QTMEM")) h->use_q_tmem = 0;
        if (flag("MPR_STATIC_TILES")) h->dynamic_tiles = 0;
        if (flag("MPR_NO_GTHR")) h->shared_thr = 0;
        if (flag("MPR_NO_FUSED_TAIL")) h->fused_tail = 0;
        if (flag("MPR_NO_COOP")) h->cooperative = 0;
        if (flag("MPR_NO_REGLIST")) h->use_reg_list = 0;
        if (flag("MPR_NO_TAIL_FLOOR")) h->tail_floor = 0;
        if (flag("MPR_NO_QCOOP")) h->q_coop = 0;
        if (flag("MPR_PDL")) h->pdl = 1;
        const char* xm = getenv("MPR_XCHG_MODE");
        if (xm) h->xchg_mode = atoi(xm);
        if (flag("MPR_NO_HYBRID")) h->use_hybrid = 0;
        const char* hb = getenv("MPR_HYBRID_MIN_B");
        if (hb && atoi(hb) >= 0) h->hybrid_min_b = atoi(hb);
        const char* tr = getenv("MPR_THR_REPLICAS");
        if (tr) h->thr_rep_log2 = tr[0] == '1' ? 0 : tr[0] == '2' ? 1 : 2;
        const char* fw = getenv("MPR_FIRST_WAIT_NS");
        if (fw) h->first_wait_ns = atoi(fw) < 0 ? -1 : (atoi(fw) > 100000 ? 100000 : atoi(fw));
        const char* dc = getenv("MPR_DEBUG_COUNTERS");
        h->dbg_counters = dc && dc[0] == '1';
        if (dc && (dc[0] == '1' || dc[0] == '2') && e == cudaSuccess) {
            e = cudaMalloc(&h->d_dbg, kDbgWords * sizeof(unsigned long long));
            if (e == cudaSuccess) e = cudaMemset(h->d_dbg, 0, kDbgWords * sizeof(unsigned long long));
        }
        const char* xt = getenv("MPR_XCHG_TIMEOUT_S");
        if (xt && atof(xt) > 0) h->xchg_timeout_ns = static_cast<unsigned long long>(atof(xt) * 1e9);
    }
    if (e != cudaSuccess) {
        fail(nullptr, MPR_ECUDA, "handle setup failed: %s", cudaGetErrorString(e));
        if (h->d_err) cudaFree(h->d_err);
        delete h;
        return MPR_ECUDA;
    }
    *out = h;
    return MPR_OK;
}

int mpr_set_exchange_timeout(mpr_handle_t h, double seconds) {
    if (!h || !(seconds > 0)) return fail(h, MPR_EINVAL, "timeout must be positive");
    h->xchg_timeout_ns = static_cast<unsigned long long>(seconds * 1e9);
    return MPR_OK;
}

int mpr_profile_begin(mpr_handle_t h, int max_launches) {
    if (!h || max_launches < 1) return fail(h, MPR_EINVAL, "bad arguments");
    DeviceGuard guard(h->device);
    while (static_cast<int>(h->prof_events.size()) < 2 * max_launches) {
        cudaEvent_t e;
        CUDA_TRY(h, cudaEventCreate(&e));
        h->prof_events.push_back(e);
    }
    h->prof_used = 0;
    return MPR_OK;
}

int mpr_profile_end(mpr_handle_t h, float* total_ms, int* n_launches) {
    if (!h || !total_ms || !n_launches) return fail(h, MPR_EINVAL, "null argument");
    DeviceGuard guard(h->device);
    const int n = h->prof_used < 0 ? 0 : h->prof_used;
    h->prof_used = -1;
    float total = 0.f;
    if (n > 0) CUDA_TRY(h, cudaEventSynchronize(h->prof_events[2 * n - 1]));
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&ms, h->prof_events[2 * i], h->prof_events[2 * i + 1]));
        total += ms;
    }
    *total_ms = total;
    *n_launches = n;
    h->prof_last_n = n;
    return MPR_OK;
}

int mpr_profile_launch_ms(mpr_handle_t h, int i, float* ms) {
    if (!h || !ms || i < 0 || i >= h->prof_last_n) return fail(h, MPR_EINVAL, "bad launch index");
    CUDA_TRY(h, cudaEventElapsedTime(ms, h->prof_events[2 * i], h->prof_events[2 * i + 1]));
    return MPR_OK;
}

int mpr_destroy(mpr_handle_t h) {
    if (!h) return MPR_OK;
    DeviceGuard guard(h->device);
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : h->io_events) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->fin_scan_done) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->fin_done) if (e) cudaEventDestroy(e);
    if (h->fin_stream) cudaStreamDestroy(h->fin_stream);
    if (h->d_finish_epoch) cudaFree(h->d_finish_epoch);
    if (h->d_err) cudaFree(h->d_err);
    if (h->d_dbg) cudaFree(h->d_dbg);
    delete h;
    return MPR_OK;
}

int mpr_debug_counters(mpr_handle_t h, uint64_t* out8) {
    if (!h || !out8) return fail(h, MPR_EINVAL, "null argument");
    memset(out8, 0, 8 * sizeof(uint64_t));
    if (!h->d_dbg) return MPR_OK;
    DeviceGuard guard(h->device);
    CUDA_TRY(h, cudaMemcpy(out8, h->d_dbg, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(h, cudaMemset(h->d_dbg, 0, 8 * sizeof(uint64_t)));
    return MPR_OK;
}

int mpr_debug_timeline(mpr_handle_t h, uint64_t* out, int n_ctas) {
    if (!h || !out || n_ctas < 1 || n_ctas > 2048) return fail(h, MPR_EINVAL, "bad argument");
    memset(out, 0, sizeof(uint64_t) * 24 * n_ctas);
    if (!h->d_dbg) return MPR_OK;
    DeviceGuard guard(h->device);
    CUDA_TRY(h, cudaMemcpy(out, h->d_dbg + 8, sizeof(uint64_t) * 24 * n_ctas, cudaMemcpyDeviceToHost));
    return MPR_OK;
}

int mpr_debug_launch_ring(mpr_handle_t h, uint64_t* out128, unsigned* next_seq) {
    if (!h || !out128) return fail(h, MPR_EINVAL, "bad argument");
    memset(out128, 0, sizeof(uint64_t) * 128);
    if (next_seq) *next_seq = h->launch_seq;
    if (!h->d_dbg) return MPR_OK;
    DeviceGuard guard(h->device);
    CUDA_TRY(h, cudaMemcpy(out128, h->d_dbg + kDbgRingOff, sizeof(uint64_t) * 128, cudaMemcpyDeviceToHost));
    CUDA_TRY(h, cudaMemset(h->d_dbg + kDbgRingOff, 0, sizeof(uint64_t) * 128));
    return MPR_OK;
}

int mpr_device_error(mpr_handle_t h, int* code) {
    if (!h || !code) return fail(h, MPR_EINVAL, "null argument");
    DeviceGuard guard(h->device);
    CUDA_TRY(h, cudaMemcpy(code, h->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (*code != 0) {
        CUDA_TRY(h, cudaMemset(h->d_err, 0, sizeof(int)));
        h->clean_ws.clear();          // a starved pipeline may have left the workspace control words dirty
    }
    return MPR_OK;
}

int mpr_bank_build(mpr_handle_t h, const void* src0, int d0, const void* src1, int d1, int src_dtype, int64_t n,
                   int normalise, uint16_t* out_bf16, float* out_bias, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (n == 0) return MPR_OK;
    if (n < 0 || !src0 || !out_bf16) return fail(h, MPR_EINVAL, "null pointer or negative n");
    if (!src1) d1 = 0;
    const int d = d0 + d1;
    if (d0 < 8 || d0 % 8 || d1 % 8 || d % 64 || d > 2048)
        return fail(h, MPR_EINVAL, "d0=%d d1=%d: parts must be multiples of 8, total a multiple of 64 and <= 2048", d0, d1);
    if (src_dtype < MPR_SRC_F32 || src_dtype > MPR_SRC_BF16) return fail(h, MPR_EINVAL, "bad src_dtype %d", src_dtype);
    if (!aligned16(src0) || !aligned16(src1) || !aligned16(out_bf16))
        return fail(h, MPR_EINVAL, "pointers must be 16-byte aligned");
    DeviceGuard guard(h->device);
    const int threads = 256, rows_per_block = threads / 32;
    long long blocks = (n + rows_per_block - 1) / rows_per_block;
    const long long cap = static_cast<long long>(h->num_sms) * 16;   // grid-stride over rows beyond this
    if (blocks > cap) blocks = cap;
    bank_build_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        src0, d0, src1, d1, src_dtype, n, normalise, out_bf16, out_bias);
    CUDA_TRY(h, cudaGetLastError());
    return MPR_OK;
}

size_t mpr_search_workspace_bytes(mpr_handle_t h, int b, int64_t n_local, int d, int kk) {
    if (!h) return 0;
    ScanPlan pl, pl2;
    if (make_plan(h, b, n_local, d, kk, &pl)) return 0;
    size_t need = ws_layout(h, pl, b, kk).total;
    // raw queries without a scratch buffer take the non-hybrid plan: size for whichever a later call may choose
    if (pl.hybrid && make_plan(h, b, n_local, d, kk, &pl2, false) == MPR_OK) {
        const size_t need2 = ws_layout(h, pl2, b, kk).total;
        if (need2 > need) need = need2;
    }
    return need;
}

int mpr_search_plan(mpr_handle_t h, int b, int64_t n_local, int d, int kk, int* n_ctas, int* n_splits, int* n_qtiles,
                    int* n_stages, int* smem_bytes) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    ScanPlan pl;
    int rc = make_plan(h, b, n_local, d, kk, &pl);
    if (rc) return rc;
    if (n_ctas) *n_ctas = pl.n_splits * pl.n_qtiles;
    if (n_splits) *n_splits = pl.n_splits;
    if (n_qtiles) *n_qtiles = pl.n_qtiles;
    if (n_stages) *n_stages = pl.n_stages;
    if (smem_bytes) *smem_bytes = static_cast<int>(pl.smem_bytes);
    return MPR_OK;
}

int mpr_plan_host(int num_sms, int b, int64_t n_local, int d, int kk, int* out16) {
    if (!out16 || num_sms < 1) return fail(nullptr, MPR_EINVAL, "bad argument");
    mpr_context ctx;                 // default tuning, no device
    ctx.num_sms = num_sms;
    ScanPlan pl;
    const int rc = make_plan(&ctx, b, n_local, d, kk, &pl);
    if (rc) {
        strncpy(g_err, ctx.err, sizeof(g_err) - 1);
        return rc;
    }
    const WsLayout wl = ws_layout(&ctx, pl, b, kk);
    const int v[16] = {pl.n_splits * pl.n_qtiles, pl.n_splits, pl.n_qtiles, pl.n_stages, static_cast<int>(pl.smem_bytes),
                       pl.q_tile, pl.sub_per_stage, pl.n_epi_groups, pl.q_tmem ? 1 : 0, pl.hybrid ? 1 : 0,
                       pl.reg_list ? 1 : 0, pl.cand_cap, pl.q_box_rows, static_cast<int>(wl.total & 0x7FFFFFFF), wl.ns,
                       1 << wl.rep_log2};
    memcpy(out16, v, sizeof(v));
    return MPR_OK;
}

int mpr_last_launch_count(mpr_handle_t h) { return h ? h->last_launches : 0; }

int mpr_workspace_invalidate(mpr_handle_t h, const void* workspace) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (!workspace) {
        h->clean_ws.clear();
        return MPR_OK;
    }
    for (size_t i = 0; i < h->clean_ws.size(); ++i)
        if (h->clean_ws[i].first == workspace) {
            h->clean_ws.erase(h->clean_ws.begin() + static_cast<long>(i));
            break;
        }
    return MPR_OK;
}

int mpr_retrieve(mpr_handle_t h, const mpr_retrieve_args* a, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    int rc = validate_retrieve(h, a);
    if (rc) return rc;
    if (a->b == 0) return MPR_OK;
    DeviceGuard guard(h->device);
    return run_step<false>(h, *a, nullptr, static_cast<cudaStream_t>(stream));
}

int mpr_retrieve_join(mpr_handle_t h, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (h->fin_seq == 0) return MPR_OK;
    DeviceGuard guard(h->device);
    CUDA_TRY(h, cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), h->fin_done[(h->fin_seq - 1) & 3], 0));
    return MPR_OK;
}

int mpr_retrieve_host(mpr_handle_t h, const mpr_retrieve_args* a, const mpr_host_io* io, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (!io) return fail(h, MPR_EINVAL, "null io block");
    int rc = validate_retrieve(h, a);
    if (rc) return rc;
    if (a->b == 0) return MPR_OK;
    DeviceGuard guard(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaStream_t st_in = io->stream_in ? static_cast<cudaStream_t>(io->stream_in) : st;
    cudaStream_t st_out = io->stream_out ? static_cast<cudaStream_t>(io->stream_out) : st;
    if ((st_in != st || st_out != st) && !h->io_events[0]) {
        for (int i = 0; i < 4; ++i) CUDA_TRY(h, cudaEventCreateWithFlags(&h->io_events[i], cudaEventDisableTiming));
    }
    // rotating pair of events: "inputs of this call are on the device" / "this call's step has run"
    cudaEvent_t ev_in = h->io_events[2 * (h->io_turn & 1)], ev_step = h->io_events[2 * (h->io_turn & 1) + 1];
    h->io_turn ^= 1;
    bool copied_in = false;
    {
        cudaStream_t st = st_in;      // the copies below go to the input stream
        static const size_t kElem[3] = {4, 2, 2};
        if (io->h_q0) {
        if (!a->q0) return fail(h, MPR_EINVAL, "h_q0 given but a->q0 (device staging) is null");
        CUDA_TRY(h, cudaMemcpyAsync(const_cast<void*>(a->q0), io->h_q0, static_cast<size_t>(a->b) * a->d0 * kElem[a->q_dtype],
                                    cudaMemcpyHostToDevice, st));
        if (io->h_q1 && a->q1)
            CUDA_TRY(h, cudaMemcpyAsync(const_cast<void*>(a->q1), io->h_q1, static_cast<size_t>(a->b) * a->d1 * kElem[a->q_dtype],
                                        cudaMemcpyHostToDevice, st));
        copied_in = true;
    }
    if (io->h_prefix_ids && a->answer_id) {
        if (!io->h_prefix_off) return fail(h, MPR_EINVAL, "h_prefix_off is null");
        if (io->n_prefix_ids > 0)
            CUDA_TRY(h, cudaMemcpyAsync(const_cast<int32_t*>(a->prefix_ids), io->h_prefix_ids,
                                        sizeof(int32_t) * static_cast<size_t>(io->n_prefix_ids), cudaMemcpyHostToDevice, st));
        CUDA_TRY(h, cudaMemcpyAsync(const_cast<int32_t*>(a->prefix_off), io->h_prefix_off,
                                    sizeof(int32_t) * (static_cast<size_t>(a->b) + 1), cudaMemcpyHostToDevice, st));
        copied_in = true;
    }
    }
    if (st_in != st && copied_in) {
        CUDA_TRY(h, cudaEventRecord(ev_in, st_in));
        CUDA_TRY(h, cudaStreamWaitEvent(st, ev_in, 0));
    }
    rc = run_step<false>(h, *a, nullptr, st);
    if (rc) return rc;
    if (io->h_out && io->d_out && io->out_bytes) {
        if (h->last_deferred) {        // the step's outputs come from the finish kernel on the library's stream
            CUDA_TRY(h, cudaStreamWaitEvent(st_out, h->fin_done[(h->fin_seq - 1) & 3], 0));
        } else if (st_out != st) {
            CUDA_TRY(h, cudaEventRecord(ev_step, st));
            CUDA_TRY(h, cudaStreamWaitEvent(st_out, ev_step, 0));
        }
        CUDA_TRY(h, cudaMemcpyAsync(io->h_out, io->d_out, io->out_bytes, cudaMemcpyDeviceToHost, st_out));
    }
    if (io->sync) CUDA_TRY(h, cudaStreamSynchronize(st_out));
    return MPR_OK;
}

int mpr_search_topk(mpr_handle_t h, const uint16_t* q, int b, const uint16_t* bank, const float* bias, int64_t n_local,
                    int64_t idx_base, int d, int kk, uint64_t* out_keys, float* out_score, int32_t* out_idx,
                    void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (b == 0) return MPR_OK;
    if (!q) return fail(h, MPR_EINVAL, "null pointer");
    mpr_retrieve_args a;
    memset(&a, 0, sizeof(a));
    a.q_bf16 = q; a.b = b; a.bank = bank; a.bias = bias; a.n_local = n_local; a.idx_base = idx_base; a.d = d; a.kk = kk;
    a.out_keys = out_keys; a.out_score = out_score; a.out_idx = out_idx;
    a.workspace = workspace; a.workspace_bytes = workspace_bytes;
    return mpr_retrieve(h, &a, stream);
}

int mpr_search_fused_supported(mpr_handle_t h, int d) { return h && d >= 64 && d <= 2048 && d % 64 == 0; }

int mpr_search_topk_fused(mpr_handle_t h, const void* src0, int d0, const void* src1, int d1, int src_dtype,
                          int normalise, int b, const uint16_t* bank, const float* bias, int64_t n_local,
                          int64_t idx_base, int kk, uint64_t* out_keys, float* out_score, int32_t* out_idx,
                          float* out_q_bias, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (b == 0) return MPR_OK;
    if (!src0) return fail(h, MPR_EINVAL, "null pointer");
    if (!src1) d1 = 0;
    if (!mpr_search_fused_supported(h, d0 + d1))
        return fail(h, MPR_EINVAL, "fused query preparation needs 64 <= D <= 2048, D %% 64 == 0; got D=%d", d0 + d1);
    mpr_retrieve_args a;
    memset(&a, 0, sizeof(a));
    a.q0 = src0; a.q1 = src1; a.d0 = d0; a.d1 = d1; a.q_dtype = src_dtype; a.normalise = normalise;
    a.b = b; a.bank = bank; a.bias = bias; a.n_local = n_local; a.idx_base = idx_base; a.d = d0 + d1; a.kk = kk;
    a.out_keys = out_keys; a.out_score = out_score; a.out_idx = out_idx; a.out_q_bias = out_q_bias;
    a.workspace = workspace; a.workspace_bytes = workspace_bytes;
    return mpr_retrieve(h, &a, stream);
}

int mpr_merge_topk(mpr_handle_t h, const uint64_t* in_keys, int n_lists, int b, int kk, uint64_t* out_keys,
                   float* out_score, int32_t* out_idx, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (b == 0) return MPR_OK;
    if (!in_keys || n_lists < 1 || b < 0) return fail(h, MPR_EINVAL, "bad arguments");
    if (kk < 1 || kk > MPR_MAX_KK) return fail(h, MPR_EINVAL, "k + skip must be in [1, %d] (got %d)", MPR_MAX_KK, kk);
    DeviceGuard guard(h->device);
    const int warps_per_block = 4;
    merge_topk_kernel<<<(b + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0,
                        static_cast<cudaStream_t>(stream)>>>(in_keys, n_lists, static_cast<long long>(b) * kk,
                                                             static_cast<long long>(kk), 1ll, b, kk, out_keys,
                                                             out_score, out_idx);
    CUDA_TRY(h, cudaGetLastError());
    return MPR_OK;
}

size_t mpr_exchange_bytes(int world, int cap) {
    if (world < 1 || world > kXchgMaxWorld || cap < 1) return 0;
    return xchg_bytes(world, cap);
}

int mpr_prompt_gather(mpr_handle_t h, const int32_t* idx, int b, int kk, int skip, const int32_t* answer_id,
                      const uint8_t* bucket_lut, const int32_t* prefix_ids, const int32_t* prefix_off,
                      const int32_t* seg_ids, const int32_t* seg_off, int use_quantifier, int pad_id, int eos_id,
                      int max_len, int out_stride, int64_t* input_ids, int64_t* attention_mask, int32_t* out_len,
                      int32_t* maj_answer, int32_t* maj_count, int32_t* bucket, int32_t* ret_answer, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (b == 0) return MPR_OK;
    if (!idx || !answer_id || !bucket_lut || !prefix_ids || !prefix_off || !seg_ids || !seg_off || !input_ids ||
        !attention_mask || !out_len || !maj_answer || !maj_count || !bucket)
        return fail(h, MPR_EINVAL, "null pointer");
    if (kk < 1 || kk > MPR_MAX_KK || skip < 0 || skip >= kk)
        return fail(h, MPR_EINVAL, "need 1 <= kk <= %d and 0 <= skip < kk (kk=%d skip=%d)", MPR_MAX_KK, kk, skip);
    if (max_len < 1 || out_stride < 1) return fail(h, MPR_EINVAL, "max_len and out_stride must be >= 1");
    DeviceGuard guard(h->device);
    PromptParams p;
    p.idx = idx; p.b = b; p.kk = kk; p.skip = skip;
    p.answer_id = answer_id; p.bucket_lut = bucket_lut;
    p.prefix_ids = prefix_ids; p.prefix_off = prefix_off; p.seg_ids = seg_ids; p.seg_off = seg_off;
    p.use_quantifier = use_quantifier; p.pad_id = pad_id; p.eos_id = eos_id;
    p.max_len = max_len; p.out_stride = out_stride;
    p.input_ids = reinterpret_cast<long long*>(input_ids);
    p.attention_mask = reinterpret_cast<long long*>(attention_mask);
    p.out_len = out_len; p.maj_answer = maj_answer; p.maj_count = maj_count; p.bucket = bucket;
    p.ret_answer = ret_answer;
    const int warps_per_block = 4;
    prompt_gather_kernel<<<(b + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0,
                           static_cast<cudaStream_t>(stream)>>>(p);
    CUDA_TRY(h, cudaGetLastError());
    return MPR_OK;
}

int mpr_embed_prompt(mpr_handle_t h, const int64_t* input_ids, const int64_t* attention_mask, int b, int len, int in_stride,
                     const void* table, int table_dtype, int vocab, int hidden, const void* image_tokens, int n_image,
                     void* out_embeds, void* out_mask, int mask_f32, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (b == 0) return MPR_OK;
    if (!input_ids || !attention_mask || !table || !out_embeds || !out_mask) return fail(h, MPR_EINVAL, "null pointer");
    if (b < 0 || len < 1 || in_stride < len || vocab < 1 || n_image < 0 || (n_image > 0 && !image_tokens))
        return fail(h, MPR_EINVAL, "bad shape (b=%d len=%d in_stride=%d vocab=%d n_image=%d)", b, len, in_stride, vocab, n_image);
    if (table_dtype < MPR_SRC_F32 || table_dtype > MPR_SRC_BF16) return fail(h, MPR_EINVAL, "bad table_dtype %d", table_dtype);
    const int esize = table_dtype == MPR_SRC_F32 ? 4 : 2;
    if (hidden < 1 || (hidden * esize) % 16 != 0) return fail(h, MPR_EINVAL, "hidden * element size must be a multiple of 16 bytes");
    if (!aligned16(table) || !aligned16(out_embeds) || !aligned16(image_tokens))
        return fail(h, MPR_EINVAL, "table, image_tokens and out_embeds must be 16-byte aligned");
    DeviceGuard guard(h->device);
    EmbedParams p;
    p.input_ids = reinterpret_cast<const long long*>(input_ids);
    p.attention_mask = reinterpret_cast<const long long*>(attention_mask);
    p.b = b; p.len = len; p.in_stride = in_stride;
    p.table = static_cast<const uint4*>(table); p.vocab = vocab; p.row_vec = hidden * esize / 16;
    p.image_tokens = static_cast<const uint4*>(image_tokens); p.n_image = n_image;
    p.out = static_cast<uint4*>(out_embeds); p.out_mask = out_mask; p.mask_f32 = mask_f32;
    p.err = h->d_err;
    const long long rows = static_cast<long long>(b) * (n_image + len);
    const int warps_per_block = 8;
    long long blocks = (rows + warps_per_block - 1) / warps_per_block;
    const long long cap = static_cast<long long>(h->num_sms) * 8;
    if (blocks > cap) blocks = cap;
    embed_prompt_kernel<<<static_cast<unsigned>(blocks), warps_per_block * 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
    CUDA_TRY(h, cudaGetLastError());
    return MPR_OK;
}

int mpr_debug_scores(mpr_handle_t h, const uint16_t* q, int b, const uint16_t* bank, const float* bias,
                     int64_t n_local, int d, float* scores, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (!q || !bank || !bias || !scores || !workspace) return fail(h, MPR_EINVAL, "null pointer");
    DeviceGuard guard(h->device);
    mpr_retrieve_args a;
    memset(&a, 0, sizeof(a));
    a.q_bf16 = q; a.b = b; a.bank = bank; a.bias = bias; a.n_local = n_local; a.d = d; a.kk = 1;
    a.workspace = workspace; a.workspace_bytes = workspace_bytes;
    return run_step<true>(h, a, scores, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
