// Kernel 2 — fused bank scan: bf16 similarity GEMM on tcgen05 (accumulators in TMEM, operands staged by
// TMA through an mbarrier ring) whose epilogue keeps a streaming per-query top-k.  The [B, N] score matrix
// of the reference (torch.cdist -> torch.argsort, /root/reference/dataset/VQAFeatureDataset.py:192-197)
// is never written.
//
// score[q, r] = <q, bank[r]> + bias[r],  bias[r] = -0.5 * ||bank[r]||^2   (so argmax score == argmin L2 distance)
//
// Orientation: queries are the MMA M dimension (one TMEM lane per query), bank rows the N dimension
// (one TMEM column per row).  Each epilogue thread therefore owns ONE query and walks its lane's columns in
// ascending row order: the common case per score is one FADD + a share of a max and a vote; only scores that
// beat the query's current k-th best are appended to a thread-private pending buffer and folded in later.
//
// Work decomposition: item = (bank split s, q-tile t); one CTA per item, blockIdx.x = s * n_qtiles + t, so the
// CTAs resident at the same time share a bank range and all but the first read of it hit L2.  The q-tile
// (<= 128 queries, <= 128 KiB) is loaded once and stays resident (tensor memory for D <= 512, else shared memory);
// only the bank streams.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..9 = two epilogue groups of four warps (warp w reads TMEM lanes 32*(w%4)...); group g takes tiles g, g+2, ...
// and keeps its own per-query lists, so a tile's epilogue may take two MMA tile-times before it stalls the pipe.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <math_constants.h>

#include "bank_build.cuh"
#include "ptx.cuh"
#include "topk_key.cuh"

namespace mpr {

constexpr int kTileRows = 128;                // bank rows per accumulator tile (UMMA N)
constexpr int kUmmaM = 128;                   // TMEM lanes = query slots per CTA
constexpr int kChunkK = 64;                   // bf16 per 128-byte swizzled row
constexpr int kStageBytes = kTileRows * 128;  // one bank K-chunk: 128 rows x 128 B = 16 KiB
constexpr int kAccBufs = 4;                   // TMEM accumulator ring: 4 x 128 columns (2 when the q-tile is in TMEM)
constexpr int kTmemCols = 512;
constexpr int kScanThreads = 320;            // warp 0 TMA, warp 1 MMA, warps 2..9 = two epilogue groups of four
constexpr int kMaxStages = 12;
constexpr int kMaxSmem = 232448;              // 227 KiB opt-in limit per CTA on sm_100
constexpr int kMaxKK = 32;
constexpr int kMaxSubPerStage = 4;           // 64-wide K sub-chunks per ring stage
constexpr int kEpiGroups = 2;                 // epilogue groups; group g owns tiles g, g+2, ... (own lists per query)
constexpr int kCandCapMax = 16;               // per-query pending-candidate slots (a flush leaves >= 8 free)

struct ScanParams {
    int b_total;       // queries in the batch
    int n_local;       // bank rows in this shard
    int n_chunks;      // D / 64
    int kk;            // list length (k + skip), 1..32
    int kk_pad;        // next power of two >= kk
    int cand_cap;      // pending-candidate slots per query (10..16)
    int q_tile;        // queries per q-tile
    int q_box_rows;    // rows of the Q TMA box (multiple of 8, >= valid rows of any q-tile)
    int n_qtiles;
    int n_splits;
    int n_tiles;       // ceil(n_local / 128)
    int n_stages;
    int n_epi_groups;  // epilogue groups actually used (1 or 2)
    int sub_per_stage; // 64-wide K sub-chunks per ring stage (1, 2 or 4): one barrier round-trip per stage
    uint32_t idx_base; // global row index of this shard's row 0
    uint64_t bank_policy;
    const float* bias;     // [n_local]
    const uint16_t* q;     // [b_total][d] bf16 queries (read directly in the TMEM-operand variant)
    int d;                 // row length
    // fused query preparation (kFuseQ): the raw CLIP halves are concatenated, optionally normalised and rounded to bf16
    // on their way into tensor memory — no separate cast kernel, no bf16 copy of the queries in HBM
    const void* qsrc0;     // [b_total][qd0]
    const void* qsrc1;     // [b_total][qd1] or nullptr
    int qd0, qd1, q_dtype, q_normalise;
    float* q_bias_out;     // [b_total] -0.5*|bf16(q)|^2 (written by split 0), or nullptr
    uint64_t* part_keys;   // [b_total][kk][n_splits * kEpiGroups]
    float* dump;           // debug: [b_total][n_local] scores, or nullptr
    int* err;              // device word that receives the code of a starved barrier
};

struct ScanSmemLayout {
    uint32_t q_off, stage_off, list_off, bias_off, bar_off, total;
};

// Per-query shared-memory row: [kk_pad sorted keys | cand_cap pending candidates | pad]; the odd stride (in
// 8-byte words) keeps the 32 lanes of a warp, each walking its own row, on distinct banks.
__host__ __device__ inline uint32_t scan_row_stride(int kk_pad, int cand_cap) {
    return (static_cast<uint32_t>(kk_pad) + cand_cap) | 1u;
}

__host__ __device__ inline ScanSmemLayout scan_smem_layout(int n_chunks, int q_box_rows, int kk_pad, int cand_cap,
                                                           int n_stages, int sub_per_stage, int n_groups) {
    ScanSmemLayout l;
    l.q_off = 0;
    l.stage_off = static_cast<uint32_t>(n_chunks) * q_box_rows * 128u;          // multiple of 1024
    l.list_off = l.stage_off + static_cast<uint32_t>(n_stages) * sub_per_stage * kStageBytes;
    l.bias_off = l.list_off + static_cast<uint32_t>(n_groups * kUmmaM) * scan_row_stride(kk_pad, cand_cap) * 8u;
    l.bar_off = l.bias_off + kAccBufs * kTileRows * 4u;
    l.total = l.bar_off + (1 + 2 * kMaxStages + 2 * kAccBufs) * 8u + 16u;
    return l;
}

// Barrier error codes (ScanParams::err)
enum : int { kErrQFull = 101, kErrEmpty = 102, kErrFull = 103, kErrTmemEmpty = 104, kErrTmemFull = 105 };

// kCluster = 2 (shared-memory q-tile variant, even number of q-tiles): the two CTAs of a cluster work on the SAME bank
// split with DIFFERENT q-tiles; each loads half of every bank K-chunk and TMA-multicasts it into both CTAs' rings, which
// halves the L2 reads per MMA.  A ring slot is refilled only after BOTH consumers have released it (multicast
// tcgen05.commit, empty count = 2).  Measured effect at B = 4096: +3 % — L2 bandwidth was not the limiter.
//
// kQTmem (D <= 512): the q-tile lives in TENSOR MEMORY (columns [0, D/2)) and is the MMA's TMEM A operand.  Only B then
// crosses the 128 B/cycle shared-memory port (an SS-mode 128x128x16 MMA alone reads 8 KiB per 64 cycles = all of it), and
// the 128 KiB the q-tile used to occupy go to the bank ring (160+ KiB in flight instead of 64).  The accumulator ring is
// then 2 x 128 columns at [256, 512).  On its own this moved nothing either; what actually bound the tensor regime
// (41 % tensor-pipe activity with no warp waiting on data) was the MMA warp's own instruction stream — see the
// warp-uniform issue loop below and the multi-sub-chunk ring stages.
// kFuseQ (with kQTmem): see ScanParams::qsrc0 — replaces /root/reference/dataset/VQAFeatureDataset.py:189-191 in-kernel.
template <bool kDump, int kCluster, bool kQTmem, bool kFuseQ = false>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_bank,
                 const ScanParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;   // SWIZZLE_128B atoms need 1024-byte alignment
    uint8_t* smem = smem_raw + (base - raw_addr);

    const ScanSmemLayout lay = scan_smem_layout(p.n_chunks, p.q_box_rows, p.kk_pad, p.cand_cap, p.n_stages, p.sub_per_stage, p.n_epi_groups);
    const uint32_t q_smem = base + lay.q_off;
    const uint32_t stage_smem = base + lay.stage_off;
    uint64_t* lists = reinterpret_cast<uint64_t*>(smem + lay.list_off);
    float* bias_s = reinterpret_cast<float*>(smem + lay.bias_off);
    const uint32_t bar_base = base + lay.bar_off;
    const uint32_t bar_q = bar_base;
    auto bar_full = [&](int s) { return bar_base + 8u + 8u * s; };
    auto bar_empty = [&](int s) { return bar_base + 8u + 8u * (kMaxStages + s); };
    auto bar_tfull = [&](int b) { return bar_base + 8u + 8u * (2 * kMaxStages + b); };
    auto bar_tempty = [&](int b) { return bar_base + 8u + 8u * (2 * kMaxStages + kAccBufs + b); };
    volatile uint32_t* tmem_slot =
        reinterpret_cast<volatile uint32_t*>(smem + lay.bar_off + (1 + 2 * kMaxStages + 2 * kAccBufs) * 8u);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int kBufs = kQTmem ? 2 : kAccBufs;               // accumulator ring depth
    constexpr uint32_t kAccCol0 = kQTmem ? 256u : 0u;          // first accumulator column

    // ---- which item is this CTA's
    const int item = blockIdx.x;
    const int split = item / p.n_qtiles;
    const int qt = item - split * p.n_qtiles;
    const int q0 = qt * p.q_tile;
    const int q_valid = min(p.q_tile, p.b_total - q0);
    const int tile_begin = static_cast<int>(static_cast<long long>(split) * p.n_tiles / p.n_splits);
    const int tile_end = static_cast<int>(static_cast<long long>(split + 1) * p.n_tiles / p.n_splits);
    const int my_tiles = tile_end - tile_begin;

    // ---- one-time setup
    if (threadIdx.x == 0) {
        ptx::mbar_init(bar_q, kQTmem ? 4 : 1);          // TMEM variant: one arrive per epilogue warp
        for (int s = 0; s < p.n_stages; ++s) {
            ptx::mbar_init(bar_full(s), 1);
            ptx::mbar_init(bar_empty(s), kCluster);      // one release per consumer CTA of the cluster
        }
        for (int b = 0; b < kBufs; ++b) {
            ptx::mbar_init(bar_tfull(b), 1);
            ptx::mbar_init(bar_tempty(b), 4);   // one arrive per epilogue warp
        }
        ptx::fence_mbar_init();
        if constexpr (!kQTmem) ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_bank);
    }
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (kCluster > 1) ptx::cluster_sync();      // peers' barriers exist before anything remote touches them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t crank = kCluster > 1 ? ptx::cluster_ctarank() : 0u;
    constexpr uint16_t kClusterMask = static_cast<uint16_t>((1u << kCluster) - 1u);

    // Producer and MMA warps run their loops with WARP-UNIFORM control flow (all 32 lanes wait on the barriers together)
    // and elect one lane only around the asynchronous instructions themselves.  Putting the whole loop under
    // `if (lane == 0)` makes every operand thread-variant for the compiler: it then wraps each UTCHMMA / UTMALDG in an
    // election loop with R2UR moves (~130 SASS instructions per K-chunk, measured: MMA issue-bound at 41 % tensor
    // activity with nothing waiting on data).
    if (warp == 0) {
        // =========================== TMA producer ===========================
        if constexpr (!kQTmem) {
            if (ptx::elect_one()) {
                // resident q-tile: one 128-byte-wide slab per K-chunk
                const uint32_t slab_bytes = static_cast<uint32_t>(p.q_box_rows) * 128u;
                ptx::mbar_arrive_expect_tx(bar_q, slab_bytes * p.n_chunks);
                for (int j = 0; j < p.n_chunks; ++j)
                    ptx::tma_load_2d(q_smem + j * slab_bytes, &tmap_q, bar_q, j * kChunkK, q0, ptx::kEvictLast);
            }
            __syncwarp();
        }
        // streamed bank: a ring stage holds up to sub_per_stage 64-wide K sub-chunks and costs ONE barrier round-trip
        const int spp = p.sub_per_stage;
        int s = 0;
        uint32_t ph = 0;
        for (int t = tile_begin; t < tile_end; ++t) {
            for (int j0 = 0; j0 < p.n_chunks; j0 += spp) {
                const int ns = min(spp, p.n_chunks - j0);
                ptx::mbar_wait(bar_empty(s), ph ^ 1u, p.err, kErrEmpty);
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(bar_full(s), ns * kStageBytes);
                    for (int u = 0; u < ns; ++u) {
                        const uint32_t dst = stage_smem + (s * spp + u) * kStageBytes;
                        if constexpr (kCluster > 1) {
                            // my 1/kCluster of the rows, delivered to every CTA of the cluster
                            constexpr uint32_t kPart = kStageBytes / kCluster;
                            ptx::tma_load_2d_multicast(dst + crank * kPart, &tmap_bank, bar_full(s), (j0 + u) * kChunkK,
                                                       t * kTileRows + crank * (kTileRows / kCluster), kClusterMask,
                                                       p.bank_policy);
                        } else {
                            ptx::tma_load_2d(dst, &tmap_bank, bar_full(s), (j0 + u) * kChunkK, t * kTileRows,
                                             p.bank_policy);
                        }
                    }
                }
                __syncwarp();
                if (++s == p.n_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(kUmmaM, kTileRows);
        // descriptor = constant high word | (address >> 4): advancing by a stage / slab / K-step is an integer add
        const uint64_t desc_hi = ptx::make_kmajor_sw128_desc(0);
        const uint32_t b_lo0 = (stage_smem >> 4) & 0x3FFFu;
        const uint32_t a_lo0 = (q_smem >> 4) & 0x3FFFu;
        const uint32_t slab_lo = static_cast<uint32_t>(p.q_box_rows) * 8u;      // slab bytes >> 4
        ptx::mbar_wait(bar_q, 0, p.err, kErrQFull);
        ptx::tc_fence_after();
        const int spp = p.sub_per_stage;
        int s = 0;
        uint32_t ph = 0;
        for (int lt = 0; lt < my_tiles; ++lt) {
            const int buf = lt & (kBufs - 1);
            const uint32_t bph = (lt / kBufs) & 1u;
            ptx::mbar_wait(bar_tempty(buf), bph ^ 1u, p.err, kErrTmemEmpty);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + kAccCol0 + buf * kTileRows;
            for (int j0 = 0; j0 < p.n_chunks; j0 += spp) {
                const int ns = min(spp, p.n_chunks - j0);
                ptx::mbar_wait(bar_full(s), ph, p.err, kErrFull);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
#pragma unroll
                    for (int u = 0; u < kMaxSubPerStage; ++u) {
                        if (u < ns) {
                            const int j = j0 + u;
                            const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(s * spp + u) * (kStageBytes >> 4);
#pragma unroll
                            for (int k = 0; k < kChunkK / 16; ++k) {
                                const uint64_t db = desc_hi | static_cast<uint64_t>(b_lo + 2u * k);   // +32 B per K-step
                                if constexpr (kQTmem) {
                                    // 16 bf16 of K = 8 TMEM columns; sub-chunk j starts at column j*32
                                    ptx::umma_bf16_ts(d_tmem, tmem_base + j * (kChunkK / 2) + k * 8, db, idesc,
                                                      (j | k) != 0 ? 1u : 0u);
                                } else {
                                    const uint64_t da = desc_hi | static_cast<uint64_t>(a_lo0 + j * slab_lo + 2u * k);
                                    ptx::umma_bf16_ss(d_tmem, da, db, idesc, (j | k) != 0 ? 1u : 0u);
                                }
                            }
                        }
                    }
                    if constexpr (kCluster > 1) ptx::umma_commit_multicast(bar_empty(s), kClusterMask);
                    else ptx::umma_commit(bar_empty(s));     // ring stage reusable once these MMAs retire
                    if (j0 + ns == p.n_chunks) ptx::umma_commit(bar_tfull(buf));   // accumulator tile complete
                }
                __syncwarp();
                if (++s == p.n_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // =========================== epilogue: streaming top-k ===========================
        // The 32 lanes of a warp are 32 independent queries, so everything here is thread-private: a lane keeps its
        // admission threshold (score of its current kk-th best) in a register, appends the rare scores that beat it
        // to its own pending buffer in shared memory, and — when some lane's buffer runs full — every lane folds
        // its own pending candidates into its own sorted list.  No cross-lane traffic except one vote per 8 scores.
        const int grp = (warp - 2) >> 2;               // epilogue group: owns tiles grp, grp + 2, ...
        const int quad = warp & 3;                     // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;              // query slot (TMEM lane) owned by this thread
        const bool valid = row < q_valid;
        const bool warp_has_work = quad * 32 < q_valid;
        const int ep_tid = ((warp - 2) & 3) * 32 + lane;   // 0..127 within the group, used to stage the bias tile
        const int kk = p.kk;
        const int kk_pad = p.kk_pad;
        uint64_t* my_list = lists + static_cast<size_t>(grp * kUmmaM + row) * scan_row_stride(kk_pad, p.cand_cap);   // [0, kk)
        uint2* my_pend = reinterpret_cast<uint2*>(my_list + kk_pad);                      // (score bits, row)
        const bool active_group = grp < p.n_epi_groups;      // an idle group owns no list memory
        if (active_group) for (int i = 0; i < kk; ++i) my_list[i] = 0ull;
        float thr = valid ? -CUDART_INF_F : CUDART_INF_F;
        int n_pend = 0;
        const int flush_at = p.cand_cap - 8;           // the next group of 8 scores must always fit

        auto flush = [&]() {
            for (int c = 0; c < n_pend; ++c) {
                const uint2 cand = my_pend[c];
                const uint64_t key = make_key(__uint_as_float(cand.x), cand.y);
                uint64_t lower = my_list[kk - 1];
                if (key > lower) {
                    // Branch-free single pass from the bottom: new[j] = old[j-1] >= key ? max(old[j], key) : old[j-1].
                    // No iteration depends on a loaded value for control flow, so the LDS/STS stream pipelines
                    // (a compare-and-break insertion loop pays one shared-memory latency per shifted element).
#pragma unroll 4
                    for (int j = kk - 1; j > 0; --j) {
                        const uint64_t upper = my_list[j - 1];
                        my_list[j] = upper >= key ? (lower > key ? lower : key) : upper;
                        lower = upper;
                    }
                    my_list[0] = lower > key ? lower : key;
                }
            }
            n_pend = 0;
            const uint64_t kth = my_list[kk - 1];
            if (valid) thr = kth == 0ull ? -CUDART_INF_F : key_score(kth);
        };

        if (kQTmem && grp == 0) {
            // This thread's query row -> its TMEM lane, columns [0, D/2): 8 bf16 (one uint4) fill 4 columns.
            const uint32_t q_taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
            const size_t qrow_idx = static_cast<size_t>(valid ? q0 + row : 0);
            if constexpr (kFuseQ) {
                auto load_q8 = [&](int col, float (&x)[8]) {      // 8 consecutive elements of [src0 | src1]
                    if (col < p.qd0) load8(p.qsrc0, p.q_dtype, qrow_idx * p.qd0 + col, x);
                    else             load8(p.qsrc1, p.q_dtype, qrow_idx * p.qd1 + (col - p.qd0), x);
                };
                float scale = 1.f;
                if (p.q_normalise && valid) {
                    float ss = 0.f;
                    for (int col = 0; col < p.d; col += 8) {
                        float x[8];
                        load_q8(col, x);
#pragma unroll
                        for (int i = 0; i < 8; ++i) ss = fmaf(x[i], x[i], ss);
                    }
                    scale = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
                }
                float rs = 0.f;
                for (int c0 = 0; c0 < p.d / 2; c0 += 32) {
                    uint32_t w[32];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        if (valid) load_q8(c0 * 2 + u * 8, x);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const __nv_bfloat16 lo = __float2bfloat16_rn(x[2 * i] * scale);
                            const __nv_bfloat16 hi = __float2bfloat16_rn(x[2 * i + 1] * scale);
                            const float flo = __bfloat162float(lo), fhi = __bfloat162float(hi);
                            rs = fmaf(flo, flo, rs);
                            rs = fmaf(fhi, fhi, rs);
                            w[4 * u + i] = static_cast<uint32_t>(__bfloat16_as_ushort(lo)) |
                                           (static_cast<uint32_t>(__bfloat16_as_ushort(hi)) << 16);
                        }
                    }
                    ptx::tmem_st_32x32b_x32(q_taddr + c0, w);
                }
                if (valid && split == 0 && p.q_bias_out) p.q_bias_out[q0 + row] = -0.5f * rs;
            } else {
                const uint4* qrow = reinterpret_cast<const uint4*>(p.q + qrow_idx * p.d);
                for (int c0 = 0; c0 < p.d / 2; c0 += 32) {
                    uint32_t w[32];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        uint4 t = make_uint4(0u, 0u, 0u, 0u);
                        if (valid) t = __ldg(qrow + c0 / 4 + u);
                        w[4 * u + 0] = t.x; w[4 * u + 1] = t.y; w[4 * u + 2] = t.z; w[4 * u + 3] = t.w;
                    }
                    ptx::tmem_st_32x32b_x32(q_taddr + c0, w);
                }
            }
            ptx::tmem_wait_st();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_q);
        }

        auto load_bias = [&](int t) -> float {
            const int r = t * kTileRows + ep_tid;
            return (r < p.n_local) ? __ldg(p.bias + r) : -CUDART_INF_F;   // -inf masks rows past the shard end
        };
        // With n_epi_groups == 1 group 1 idles: it only writes its (empty) lists.  Used when list maintenance dominates
        // and both groups' active warps would share one SM sub-partition anyway (few queries, large k).
        const int n_groups = p.n_epi_groups;
        const int first_tile = grp < n_groups ? grp : my_tiles;
        float next_bias = first_tile < my_tiles ? load_bias(tile_begin + first_tile) : 0.f;

        for (int lt = first_tile; lt < my_tiles; lt += n_groups) {
            const int t = tile_begin + lt;
            const int buf = lt & (kBufs - 1);
            const uint32_t bph = (lt / kBufs) & 1u;
            // bias tiles are double-buffered per group: a fast warp may stage tile lt+2 while a slow one still reads lt
            float* bias_tile = bias_s + (grp * 2 + ((lt / n_groups) & 1)) * kTileRows;
            bias_tile[ep_tid] = next_bias;
            ptx::named_bar_sync(1 + grp, 128);
            if (lt + n_groups < my_tiles) next_bias = load_bias(t + n_groups);

            ptx::mbar_wait(bar_tfull(buf), bph, p.err, kErrTmemFull);
            ptx::tc_fence_after();

            if (warp_has_work) {
                const uint32_t row_base = static_cast<uint32_t>(t) * kTileRows;
#pragma unroll 1
                for (int c0 = 0; c0 < kTileRows; c0 += 32) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + kAccCol0 +
                                                buf * kTileRows + c0, v);
                    ptx::tmem_wait_ld();
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float s[8];
                        const float4 b0 = *reinterpret_cast<const float4*>(bias_tile + c0 + g * 8);
                        const float4 b1 = *reinterpret_cast<const float4*>(bias_tile + c0 + g * 8 + 4);
                        s[0] = __uint_as_float(v[g * 8 + 0]) + b0.x;
                        s[1] = __uint_as_float(v[g * 8 + 1]) + b0.y;
                        s[2] = __uint_as_float(v[g * 8 + 2]) + b0.z;
                        s[3] = __uint_as_float(v[g * 8 + 3]) + b0.w;
                        s[4] = __uint_as_float(v[g * 8 + 4]) + b1.x;
                        s[5] = __uint_as_float(v[g * 8 + 5]) + b1.y;
                        s[6] = __uint_as_float(v[g * 8 + 6]) + b1.z;
                        s[7] = __uint_as_float(v[g * 8 + 7]) + b1.w;
                        if constexpr (kDump) {
                            if (valid) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const uint32_t r = row_base + c0 + g * 8 + i;
                                    if (r < static_cast<uint32_t>(p.n_local))
                                        p.dump[static_cast<size_t>(q0 + row) * p.n_local + r] = s[i];
                                }
                            }
                        }
                        const float m = fmaxf(fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])),
                                              fmaxf(fmaxf(s[4], s[5]), fmaxf(s[6], s[7])));
                        if (__any_sync(kFullMask, m > thr)) {
                            // Rows arrive in ascending order, so a later row can only displace the kk-th best with a
                            // STRICTLY higher score: `>` is the exact admission test for (score desc, row asc).
                            const uint32_t gidx = p.idx_base + row_base + c0 + g * 8;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (s[i] > thr) {
                                    my_pend[n_pend] = make_uint2(__float_as_uint(s[i]), gidx + i);
                                    ++n_pend;
                                }
                            }
                            if (__any_sync(kFullMask, n_pend > flush_at)) flush();
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_tempty(buf));
        }

        // partial result of this (split, group, q-tile): part_keys[q0 + row][rank][split * 2 + grp] — rank-major per
        // query, so the merge kernel's walk over all lists' rank-i candidates is one contiguous stream
        if (warp_has_work) {
            if (active_group) flush();
            if (valid) {
                const size_t n_lists = static_cast<size_t>(p.n_splits) * kEpiGroups;
                uint64_t* dst = p.part_keys + static_cast<size_t>(q0 + row) * kk * n_lists + split * kEpiGroups + grp;
                for (int i = 0; i < kk; ++i) dst[static_cast<size_t>(i) * n_lists] = active_group ? my_list[i] : 0ull;
            }
        }
    }

    // ---- teardown
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (kCluster > 1) ptx::cluster_sync();      // no CTA leaves while a peer may still signal its barriers
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace mpr
