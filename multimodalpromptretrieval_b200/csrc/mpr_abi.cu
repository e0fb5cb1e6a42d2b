// C-ABI entry points (include/mpr_b200.h): argument validation, launch planning, TMA descriptor encoding and
// kernel launches.  No device synchronisation, no persistent device allocations beyond one error word.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/mpr_b200.h"
#include "bank_build.cuh"
#include "exchange.cuh"
#include "merge_topk.cuh"
#include "prompt_gather.cuh"
#include "scan_topk.cuh"

using namespace mpr;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct mpr_context {
    int device = -1;
    int num_sms = 0;
    int* d_err = nullptr;
    PFN_encodeTiled encode = nullptr;
    std::vector<cudaEvent_t> prof_events;   // start/stop pairs, used only between mpr_profile_begin/end
    int cand_cap_override = 0;              // MPR_CAND_CAP=10..16 forces the pending-buffer depth
    int epi_groups = 2;                     // MPR_EPI_GROUPS=1 forces a single epilogue group
    int stage_subs = 4;                     // max 64-wide K sub-chunks per ring stage (MPR_STAGE_SUBS=1|2|4)
    int use_q_tmem = 1;                     // q-tile as TMEM A operand when D <= 512 (MPR_NO_QTMEM=1 disables)
    int use_cluster = 1;                    // CTA-pair TMA multicast in the tensor-bound regime (MPR_NO_CLUSTER=1 disables)
    int prof_used = -1;                     // -1 = profiling off
    int prof_last_n = 0;                    // launches recorded by the last begin/end pair
    char err[512] = {0};
};

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

// ------------------------------------------------------------------------------------------------ planning
struct ScanPlan {
    int n_chunks, q_tile, q_box_rows, n_qtiles, n_splits, n_tiles, n_stages, kk_pad, cand_cap, sub_per_stage, n_epi_groups;
    bool q_tmem;       // q-tile in tensor memory (TMEM A operand) instead of shared memory
    uint32_t smem_bytes;
};

static int make_plan(mpr_context* h, int b, int64_t n_local, int d, int kk, ScanPlan* pl) {
    if (b < 1) return fail(h, MPR_EINVAL, "b must be >= 1 (got %d)", b);
    if (n_local < 1 || n_local >= (1ll << 31) - kTileRows)
        return fail(h, MPR_EINVAL, "n_local out of range (got %lld)", static_cast<long long>(n_local));
    if (d < 64 || d > 4096 || d % 64 != 0) return fail(h, MPR_EINVAL, "d must be a multiple of 64 in [64, 4096] (got %d)", d);
    if (kk < 1 || kk > MPR_MAX_KK) return fail(h, MPR_EINVAL, "k + skip must be in [1, %d] (got %d)", MPR_MAX_KK, kk);

    pl->n_chunks = d / kChunkK;
    pl->kk_pad = pow2_ceil(kk);
    pl->n_tiles = static_cast<int>((n_local + kTileRows - 1) / kTileRows);
    // D <= 512: the q-tile (128 x D bf16) fits 256 TMEM columns next to two 128-column accumulators
    pl->q_tmem = h->use_q_tmem && d <= 512;
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
        if (pl->q_tmem) pl->q_box_rows = 0;       // nothing of Q in shared memory
        // List maintenance dominates for k+s >= 16: one epilogue group (one list per query instead of two, ~1.8x less
        // insert work); otherwise two groups.
        // The second group's lists must also not starve the bank ring (SS mode keeps 64-128 KiB of Q in shared memory).
        // (deeper pending buffers were measured to HURT: 32 slots -> +25 % at k+s = 16/32, because the admission
        // threshold only moves at a flush and a stale threshold admits many more candidates)
        const int caps1[] = {16, 14, 12, 10, 10}, caps2[] = {16, 14, 12, 10};
        auto units_for = [&](int cap, int groups) {   // 16 KiB ring units left beside the resident q-tile and the lists
            const ScanSmemLayout fixed = scan_smem_layout(pl->n_chunks, pl->q_box_rows, pl->kk_pad, cap, 0, 1, groups);
            return (kMaxSmem - 1024 - static_cast<int>(fixed.total)) / kStageBytes;
        };
        pl->n_epi_groups = (h->epi_groups == 1 || pl->kk_pad >= 16 || units_for(16, 2) < 6) ? 1 : 2;
        const int* caps = pl->n_epi_groups == 1 ? caps1 : caps2;
        const int n_caps = pl->n_epi_groups == 1 ? 5 : 4;
        pl->cand_cap = caps[n_caps - 1];
        stages = units_for(pl->cand_cap, pl->n_epi_groups);
        for (int want : {6, 3}) {
            bool found = false;
            for (int c = 0; c < n_caps && !found; ++c)
                if (units_for(caps[c], pl->n_epi_groups) >= want) {
                    pl->cand_cap = caps[c];
                    stages = units_for(caps[c], pl->n_epi_groups);
                    found = true;
                }
            if (found) break;
        }
        if (h->cand_cap_override >= 10 && h->cand_cap_override <= kCandCapMax && units_for(h->cand_cap_override, pl->n_epi_groups) >= 3) {
            pl->cand_cap = h->cand_cap_override;      // tuning knob (MPR_CAND_CAP)
            stages = units_for(pl->cand_cap, pl->n_epi_groups);
        }
        if (stages >= 3 || q_tile_max <= 32 || pl->q_tmem || pl->q_box_rows < q_tile_max) break;
        q_tile_max >>= 1;
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
    pl->smem_bytes = scan_smem_layout(pl->n_chunks, pl->q_box_rows, pl->kk_pad, pl->cand_cap, stages, pl->sub_per_stage, pl->n_epi_groups).total + 1024u;
    return MPR_OK;
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

struct FusedQ {            // raw query halves for the fused-cast variant (src0 == nullptr: q is already bf16)
    const void* src0 = nullptr;
    const void* src1 = nullptr;
    int d0 = 0, d1 = 0, dtype = 0, normalise = 0;
    float* q_bias_out = nullptr;
};

template <bool kDump>
static int launch_scan(mpr_context* h, const ScanPlan& pl, const uint16_t* q, int b, const uint16_t* bank,
                       const float* bias, int64_t n_local, int64_t idx_base, int d, int kk, uint64_t* part_keys,
                       float* dump, cudaStream_t st, const FusedQ& fq = FusedQ()) {
    CUtensorMap tq, tb;
    int rc = MPR_OK;
    if (pl.q_tmem) memset(&tq, 0, sizeof(tq));       // the TMEM variant never touches the Q tensor map
    else rc = encode_2d(h, &tq, q, static_cast<uint64_t>(b), static_cast<uint64_t>(d), pl.q_box_rows);
    if (rc) return rc;
    // tensor-bound regime with an even number of q-tiles: CTA pairs share each bank chunk by TMA multicast
    const bool pair = !kDump && !pl.q_tmem && h->use_cluster && pl.n_qtiles >= 2 && pl.n_qtiles % 2 == 0;
    rc = encode_2d(h, &tb, bank, static_cast<uint64_t>(n_local), static_cast<uint64_t>(d), pair ? kTileRows / 2 : kTileRows);
    if (rc) return rc;

    ScanParams p;
    p.b_total = b;
    p.n_local = static_cast<int>(n_local);
    p.n_chunks = pl.n_chunks;
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
    p.idx_base = static_cast<uint32_t>(idx_base);
    p.bank_policy = pl.n_qtiles == 1 ? ptx::kEvictFirst : ptx::kEvictNormal;
    p.bias = bias;
    p.q = q;
    p.d = d;
    p.qsrc0 = fq.src0;
    p.qsrc1 = fq.src1;
    p.qd0 = fq.d0;
    p.qd1 = fq.d1;
    p.q_dtype = fq.dtype;
    p.q_normalise = fq.normalise;
    p.q_bias_out = fq.q_bias_out;
    p.part_keys = part_keys;
    p.dump = dump;
    p.err = h->d_err;

    const dim3 grid(pl.n_splits * pl.n_qtiles);
    const bool prof = h->prof_used >= 0 && 2 * (h->prof_used + 1) <= static_cast<int>(h->prof_events.size());
    if (prof) CUDA_TRY(h, cudaEventRecord(h->prof_events[2 * h->prof_used], st));
    if (pair) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(kScanThreads);
        cfg.dynamicSmemBytes = pl.smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CUDA_TRY(h, cudaLaunchKernelEx(&cfg, scan_topk_kernel<false, 2, false>, tq, tb, p));
    } else if (pl.q_tmem && fq.src0) {
        scan_topk_kernel<false, 1, true, true><<<grid, kScanThreads, pl.smem_bytes, st>>>(tq, tb, p);
    } else if (pl.q_tmem) {
        scan_topk_kernel<kDump, 1, true><<<grid, kScanThreads, pl.smem_bytes, st>>>(tq, tb, p);
    } else {
        scan_topk_kernel<kDump, 1, false><<<grid, kScanThreads, pl.smem_bytes, st>>>(tq, tb, p);
    }
    CUDA_TRY(h, cudaGetLastError());
    if (prof) {
        CUDA_TRY(h, cudaEventRecord(h->prof_events[2 * h->prof_used + 1], st));
        ++h->prof_used;
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
    CUDA_TRY(nullptr, cudaSetDevice(device));
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
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(scan_topk_kernel<false, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(scan_topk_kernel<true, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(scan_topk_kernel<false, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(scan_topk_kernel<false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(scan_topk_kernel<true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(scan_topk_kernel<false, 1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    {
        const char* nc = getenv("MPR_NO_CLUSTER");
        if (nc && nc[0] == '1') h->use_cluster = 0;
        const char* cc = getenv("MPR_CAND_CAP");
        if (cc) h->cand_cap_override = atoi(cc);
        const char* eg = getenv("MPR_EPI_GROUPS");
        if (eg && eg[0] == '1') h->epi_groups = 1;
        const char* ss = getenv("MPR_STAGE_SUBS");
        if (ss && (ss[0] == '1' || ss[0] == '2' || ss[0] == '4')) h->stage_subs = ss[0] - '0';
        const char* nq = getenv("MPR_NO_QTMEM");
        if (nq && nq[0] == '1') h->use_q_tmem = 0;
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

int mpr_profile_begin(mpr_handle_t h, int max_launches) {
    if (!h || max_launches < 1) return fail(h, MPR_EINVAL, "bad arguments");
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
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    if (h->d_err) cudaFree(h->d_err);
    delete h;
    return MPR_OK;
}

int mpr_device_error(mpr_handle_t h, int* code) {
    if (!h || !code) return fail(h, MPR_EINVAL, "null argument");
    CUDA_TRY(h, cudaMemcpy(code, h->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (*code != 0) CUDA_TRY(h, cudaMemset(h->d_err, 0, sizeof(int)));
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
    ScanPlan pl;
    if (make_plan(h, b, n_local, d, kk, &pl)) return 0;
    return static_cast<size_t>(pl.n_splits) * kEpiGroups * b * kk * sizeof(uint64_t);
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

int mpr_search_topk(mpr_handle_t h, const uint16_t* q, int b, const uint16_t* bank, const float* bias, int64_t n_local,
                    int64_t idx_base, int d, int kk, uint64_t* out_keys, float* out_score, int32_t* out_idx,
                    void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (b == 0) return MPR_OK;
    if (!q || !bank || !bias || !workspace) return fail(h, MPR_EINVAL, "null pointer");
    if (!aligned16(q) || !aligned16(bank) || !aligned16(workspace))
        return fail(h, MPR_EINVAL, "q, bank and workspace must be 16-byte aligned");
    if (idx_base < 0 || idx_base + n_local >= 0xFFFFFFFFll) return fail(h, MPR_EINVAL, "global row index exceeds 32 bits");
    ScanPlan pl;
    int rc = make_plan(h, b, n_local, d, kk, &pl);
    if (rc) return rc;
    const size_t need = static_cast<size_t>(pl.n_splits) * kEpiGroups * b * kk * sizeof(uint64_t);
    if (workspace_bytes < need)
        return fail(h, MPR_EWORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint64_t* part = static_cast<uint64_t*>(workspace);
    rc = launch_scan<false>(h, pl, q, b, bank, bias, n_local, idx_base, d, kk, part, nullptr, st);
    if (rc) return rc;
    const int warps_per_block = 4;
    merge_topk_kernel<<<(b + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(
        part, pl.n_splits * kEpiGroups, 1ll, static_cast<long long>(kk) * pl.n_splits * kEpiGroups,
        static_cast<long long>(pl.n_splits) * kEpiGroups, b, kk, out_keys, out_score, out_idx);
    CUDA_TRY(h, cudaGetLastError());
    return MPR_OK;
}

int mpr_search_fused_supported(mpr_handle_t h, int d) { return h && h->use_q_tmem && d >= 64 && d <= 512 && d % 64 == 0; }

int mpr_search_topk_fused(mpr_handle_t h, const void* src0, int d0, const void* src1, int d1, int src_dtype,
                          int normalise, int b, const uint16_t* bank, const float* bias, int64_t n_local,
                          int64_t idx_base, int kk, uint64_t* out_keys, float* out_score, int32_t* out_idx,
                          float* out_q_bias, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (b == 0) return MPR_OK;
    if (!src1) d1 = 0;
    const int d = d0 + d1;
    if (!src0 || !bank || !bias || !workspace) return fail(h, MPR_EINVAL, "null pointer");
    if (d0 < 8 || d0 % 8 || d1 % 8) return fail(h, MPR_EINVAL, "query halves must be multiples of 8 wide (d0=%d d1=%d)", d0, d1);
    if (!mpr_search_fused_supported(h, d))
        return fail(h, MPR_EINVAL, "fused query preparation needs the tensor-memory q-tile (64 <= D <= 512, D %% 64 == 0); got D=%d", d);
    if (src_dtype < MPR_SRC_F32 || src_dtype > MPR_SRC_BF16) return fail(h, MPR_EINVAL, "bad src_dtype %d", src_dtype);
    if (!aligned16(src0) || !aligned16(src1) || !aligned16(bank) || !aligned16(workspace))
        return fail(h, MPR_EINVAL, "pointers must be 16-byte aligned");
    if (idx_base < 0 || idx_base + n_local >= 0xFFFFFFFFll) return fail(h, MPR_EINVAL, "global row index exceeds 32 bits");
    ScanPlan pl;
    int rc = make_plan(h, b, n_local, d, kk, &pl);
    if (rc) return rc;
    const size_t need = static_cast<size_t>(pl.n_splits) * kEpiGroups * b * kk * sizeof(uint64_t);
    if (workspace_bytes < need) return fail(h, MPR_EWORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint64_t* part = static_cast<uint64_t*>(workspace);
    FusedQ fq;
    fq.src0 = src0; fq.src1 = src1; fq.d0 = d0; fq.d1 = d1; fq.dtype = src_dtype; fq.normalise = normalise;
    fq.q_bias_out = out_q_bias;
    rc = launch_scan<false>(h, pl, nullptr, b, bank, bias, n_local, idx_base, d, kk, part, nullptr, st, fq);
    if (rc) return rc;
    const int warps_per_block = 4;
    merge_topk_kernel<<<(b + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(
        part, pl.n_splits * kEpiGroups, 1ll, static_cast<long long>(kk) * pl.n_splits * kEpiGroups,
        static_cast<long long>(pl.n_splits) * kEpiGroups, b, kk, out_keys, out_score, out_idx);
    CUDA_TRY(h, cudaGetLastError());
    return MPR_OK;
}

int mpr_merge_topk(mpr_handle_t h, const uint64_t* in_keys, int n_lists, int b, int kk, uint64_t* out_keys,
                   float* out_score, int32_t* out_idx, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (b == 0) return MPR_OK;
    if (!in_keys || n_lists < 1 || b < 0) return fail(h, MPR_EINVAL, "bad arguments");
    if (kk < 1 || kk > MPR_MAX_KK) return fail(h, MPR_EINVAL, "k + skip must be in [1, %d] (got %d)", MPR_MAX_KK, kk);
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

int mpr_exchange_push(mpr_handle_t h, const uint64_t* local_keys, int b, int kk, int rank, int world,
                      void* const* peer_bufs, int cap, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (!local_keys || !peer_bufs) return fail(h, MPR_EINVAL, "null pointer");
    if (world < 1 || world > kXchgMaxWorld || rank < 0 || rank >= world)
        return fail(h, MPR_EINVAL, "bad rank/world %d/%d (max world %d)", rank, world, kXchgMaxWorld);
    if (b < 1 || kk < 1 || kk > MPR_MAX_KK || static_cast<long long>(b) * kk > cap)
        return fail(h, MPR_EINVAL, "b*kk = %lld exceeds the exchange capacity %d", static_cast<long long>(b) * kk, cap);
    XchgPeers peers;
    for (int r = 0; r < kXchgMaxWorld; ++r) peers.buf[r] = r < world ? static_cast<unsigned char*>(peer_bufs[r]) : nullptr;
    for (int r = 0; r < world; ++r)
        if (!peers.buf[r] || !aligned16(peers.buf[r])) return fail(h, MPR_EINVAL, "peer buffer %d is null or unaligned", r);
    xchg_push_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(local_keys, b * kk, rank, world, cap, peers);
    CUDA_TRY(h, cudaGetLastError());
    return MPR_OK;
}

int mpr_exchange_merge(mpr_handle_t h, void* my_buf, int world, int cap, int b, int kk, uint64_t* out_keys,
                       float* out_score, int32_t* out_idx, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (!my_buf || !aligned16(my_buf)) return fail(h, MPR_EINVAL, "exchange buffer is null or unaligned");
    if (world < 1 || world > kXchgMaxWorld) return fail(h, MPR_EINVAL, "bad world %d", world);
    if (b < 1 || kk < 1 || kk > MPR_MAX_KK || static_cast<long long>(b) * kk > cap)
        return fail(h, MPR_EINVAL, "b*kk = %lld exceeds the exchange capacity %d", static_cast<long long>(b) * kk, cap);
    const int warps_per_block = 4;
    xchg_merge_kernel<<<(b + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0,
                        static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned char*>(my_buf), world, cap, b, kk,
                                                             out_keys, out_score, out_idx, h->d_err);
    CUDA_TRY(h, cudaGetLastError());
    return MPR_OK;
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

int mpr_debug_scores(mpr_handle_t h, const uint16_t* q, int b, const uint16_t* bank, const float* bias,
                     int64_t n_local, int d, float* scores, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return fail(nullptr, MPR_EINVAL, "null handle");
    if (!q || !bank || !bias || !scores || !workspace) return fail(h, MPR_EINVAL, "null pointer");
    ScanPlan pl;
    int rc = make_plan(h, b, n_local, d, 1, &pl);
    if (rc) return rc;
    const size_t need = static_cast<size_t>(pl.n_splits) * kEpiGroups * b * sizeof(uint64_t);
    if (workspace_bytes < need) return fail(h, MPR_EWORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
    return launch_scan<true>(h, pl, q, b, bank, bias, n_local, 0, d, 1, static_cast<uint64_t*>(workspace), scores,
                             static_cast<cudaStream_t>(stream));
}

}  // extern "C"
