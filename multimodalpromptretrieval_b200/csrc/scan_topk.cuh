// Kernel 2 — fused bank scan: bf16 similarity GEMM on tcgen05 (accumulators in TMEM, operands staged by
// TMA through an mbarrier ring) whose epilogue keeps a streaming per-query top-k, and whose tail — when the grid is a
// single wave — finishes the whole retrieval step in the same launch (tail.cuh).  The [B, N] score matrix of the
// reference (torch.cdist -> torch.argsort, /root/reference/dataset/VQAFeatureDataset.py:192-197) is never written.
//
// score[q, r] = <q, bank[r]> + bias[r],  bias[r] = -0.5 * ||bank[r]||^2   (so argmax score == argmin L2 distance)
//
// Orientation: queries are the MMA M dimension (one TMEM lane per query), bank rows the N dimension
// (one TMEM column per row).  Each epilogue thread therefore owns ONE query and walks its lane's columns in
// ascending row order: the common case per score is one FADD + a share of a max and a vote; only scores that
// beat the query's current admission threshold are appended to a thread-private pending buffer and folded in later.
//
// Work decomposition: the CTAs of one q-tile (<= 128 queries, loaded once, resident in tensor memory for D <= 512, else
// in shared memory) pull 128-row bank tiles from a shared atomic counter (dynamic scheduling: HBM speed differs by a
// few percent between SMs, a static split leaves the fast ones idle at the end); blockIdx.x = s * n_qtiles + t, so the
// CTAs resident at the same time walk the bank together and all but the first read of a tile hit L2.
//
// Shared admission thresholds: every (CTA, epilogue group) list publishes its best score with atomicMax into one of
// ns >= kk slots per query.  The lists cover disjoint rows, so the MINIMUM over a query's ns slots is a lower bound on
// its global kk-th best score: ns distinct rows are known to score at least that much, and anything strictly below it
// can be dropped by every CTA without ever entering a list.  This replaces each CTA's slowly converging private
// threshold (k-th best of ITS rows) by a near-global one, which is what keeps list maintenance off the critical path
// for large k and for small per-GPU shards.  A list's FIRST tile is walked twice: pass 1 only publishes the tile's
// maximum, then the warp waits (bounded) for every slot of its queries to be filled, and pass 2 admits against that
// bound — a blind first tile cost 15 us (k = 5) to 30 us (k + skip = 32) per launch.
//
// Warp roles (320 threads): warp 0 = TMA producer + tile scheduler, warp 1 = TMEM allocator + MMA issuer (one elected
// lane), warps 2..9 = two epilogue groups of four warps (warp w reads TMEM lanes 32*(w%4)...); group g takes this CTA's
// tiles g, g+2, ... and keeps its own per-query lists, so a tile's epilogue may take two MMA tile-times before it stalls
// the pipe.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <math_constants.h>
#include <type_traits>

#include "bank_build.cuh"
#include "ptx.cuh"
#include "tail.cuh"
#include "topk_key.cuh"

namespace mpr {

constexpr int kTileRows = 128;                // bank rows per accumulator tile (UMMA N)
constexpr int kUmmaM = 128;                   // TMEM lanes = query slots per CTA
constexpr int kChunkK = 64;                   // bf16 per 128-byte swizzled row
constexpr int kStageBytes = kTileRows * 128;  // one bank K-chunk: 128 rows x 128 B = 16 KiB
constexpr int kAccBufs = 4;                   // TMEM accumulator ring: 4 x 128 columns (2 when the q-tile is in TMEM)
constexpr int kTmemCols = 512;
constexpr int kScanThreads = 320;            // warp 0 TMA, warp 1 MMA, warps 2..9 = two epilogue groups of four
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kMaxStages = 12;
constexpr int kMaxSmem = 232448;              // 227 KiB opt-in limit per CTA on sm_100
constexpr int kMaxKK = 32;
constexpr int kMaxSubPerStage = 4;           // 64-wide K sub-chunks per ring stage
constexpr int kEpiGroups = 2;                 // epilogue groups; group g owns tiles g, g+2, ... (own lists per query)
constexpr int kCandCapMax = 16;               // per-query pending-candidate slots (a flush leaves >= 8 free)
constexpr int kTileRing = 64;                 // scheduler -> consumers tile-id ring (entries); >= the producer's lead
constexpr uint32_t kTileEnd = 0xFFFFFFFFu;
constexpr int kQTmemChunks = 8;               // K-chunks of a q-tile that fit tensor memory (256 columns)

struct ScanParams {
    int b_total;       // queries in the batch
    int n_local;       // bank rows in this shard
    int n_chunks;      // D / 64
    int n_q_smem;      // K-chunks of the q-tile that live in shared memory: all of them without the TMEM q-tile, none
                       // with it for D <= 512, and chunks 8.. of a HYBRID q-tile (512 < D <= 1024: the first 512 dims in
                       // tensor memory, the rest in shared memory, so that 128 queries x 1024 dims stay resident at once)
    int kk;            // list length (k + skip), 1..32
    int kk_pad;        // next power of two >= kk
    int cand_cap;      // pending-candidate slots per query (10..16)
    int q_tile;        // queries per q-tile
    int q_box_rows;    // rows of the Q TMA box (multiple of 8, >= valid rows of any q-tile)
    int n_qtiles;
    int n_splits;      // CTAs per q-tile
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
    // on their way into tensor / shared memory — no separate cast kernel, no bf16 copy of the queries in HBM
    const void* qsrc0;     // [b_total][qd0]
    const void* qsrc1;     // [b_total][qd1] or nullptr
    int qd0, qd1, q_dtype, q_normalise;
    float* q_bias_out;     // [b_total] -0.5*|bf16(q)|^2 (written by split 0), or nullptr
    uint64_t* part_keys;   // [b_total][n_splits * kEpiGroups][kk], each list unordered
    uint32_t* tile_ctr;    // [n_qtiles] dynamic tile scheduler (nullptr: static contiguous ranges)
    uint32_t* gthr;        // [ns][b_total] shared admission thresholds, ordered-u32 scores (nullptr: off); slot-major so
                           // that the 32 lanes (= 32 queries) of a warp touch 4 sectors per slot, not 32
    int ns;                // threshold slots per query (multiple of 4, >= kk)
    int thr_rep_log2;      // every slot is kept as 2^r replicas (word slot * R + replica; a list feeds one of them, a reader
                           // takes the maximum): the first publish of ~300 lists lands at the same instant, and with one
                           // word per slot each 128-byte line took ~37 atomic requests in a row.  ns << r <= 32
    int fused_tail;        // 1: grid barrier + tail.cuh in this launch (grid must be co-resident)
    int q_coop;            // TMEM q-tile filled warp-cooperatively: coalesced loads -> swizzled scratch in the (not yet used)
                           // list memory -> each thread's row -> tcgen05.st (needs >= 32 KiB of list memory, no normalise)
    int first_wait_ns;     // first tile of a list: how long to wait for every threshold slot of the query to be filled
                           // after the tile's maximum was published (0: a single look; < 0: legacy blind-chunk start)
    unsigned long long* dbg;  // debug counters (MPR_DEBUG_COUNTERS=1) or nullptr: [0] candidates admitted, [1] warp flushes,
                           // [2] slow-path 8-groups (per warp), [3] list replacements, [4] warp-tiles, [5] threshold refreshes that found a bound
    unsigned long long* dbg_ts;  // debug timeline (MPR_DEBUG_COUNTERS=1|2) or nullptr: [16 * cta + event]
    unsigned long long* dbg_ring;  // debug: per-launch [first CTA entry (stored as 2^63 - t), last CTA exit] x 64, or nullptr
    unsigned launch_seq;   // host-side launch counter (selects the ring entry)
    float* dump;           // debug: [b_total][n_local] scores, or nullptr
    int* err;              // device word that receives the code of a starved barrier
};

struct ScanSmemLayout {
    uint32_t q_off, stage_off, list_off, bias_off, ring_off, scr_off, bar_off, total;
};

// Per-query shared-memory row: [kk_pad sorted keys | cand_cap pending candidates | pad]; the odd stride (in
// 8-byte words) keeps the 32 lanes of a warp, each walking its own row, on distinct banks.
__host__ __device__ inline uint32_t scan_row_stride(int kk_pad, int cand_cap) {
    return (static_cast<uint32_t>(kk_pad) + cand_cap) | 1u;
}

__host__ __device__ inline ScanSmemLayout scan_smem_layout(int n_chunks /* of the q-tile in shared memory */, int q_box_rows, int kk_pad, int cand_cap,
                                                           int n_stages, int sub_per_stage, int n_groups) {
    ScanSmemLayout l;
    l.q_off = 0;
    l.stage_off = static_cast<uint32_t>(n_chunks) * q_box_rows * 128u;          // multiple of 1024
    l.list_off = l.stage_off + static_cast<uint32_t>(n_stages) * sub_per_stage * kStageBytes;
    l.bias_off = l.list_off + static_cast<uint32_t>(n_groups * kUmmaM) * scan_row_stride(kk_pad, cand_cap) * 8u;
    l.ring_off = l.bias_off + kAccBufs * kTileRows * 4u;
    l.scr_off = l.ring_off + kTileRing * 8u;
    l.bar_off = l.scr_off + 3 * kUmmaM * 4u;
    l.total = l.bar_off + (2 + 2 * kMaxStages + 2 * kAccBufs) * 8u + 16u;
    return l;
}

// Barrier error codes (ScanParams::err)
enum : int { kErrQFull = 101, kErrEmpty = 102, kErrFull = 103, kErrTmemEmpty = 104, kErrTmemFull = 105, kErrRing = 106,
             kErrGridBarrier = 107 };

// One row of the query batch: concat + (normalise) + bf16 rounding exactly as kernel 1 does it (same per-lane
// accumulation order and the same warp reduction, so `-0.5 * sum` equals kernel 1's bias bit for bit), written as 16-byte
// units into a K-major SWIZZLE_128B shared-memory q-tile (slab j = K-chunk j, row r at r*128, unit c at (c ^ (r&7))*16).
// `col_begin` (a multiple of 256): only columns >= col_begin are read and written, slab 0 = the chunk at col_begin
// (the shared-memory half of a hybrid q-tile); normalisation needs the whole row and is not available with it.
__device__ __forceinline__ float prep_query_row_to_smem(const ScanParams& p, long long src_row, bool real, uint8_t* q_tile,
                                                        uint32_t slab_bytes, int r, int lane, int col_begin = 0) {
    const int steps = (p.d + 255) / 256;
    float x[kBuildMaxSteps][8];
    float ss = 0.f;
#pragma unroll
    for (int s = 0; s < kBuildMaxSteps; ++s) {
        const int col = s * 256 + lane * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) x[s][i] = 0.f;
        if (s < steps && col < p.d && col >= col_begin && real) {
            if (col < p.qd0) load8(p.qsrc0, p.q_dtype, static_cast<size_t>(src_row) * p.qd0 + col, x[s]);
            else             load8(p.qsrc1, p.q_dtype, static_cast<size_t>(src_row) * p.qd1 + (col - p.qd0), x[s]);
#pragma unroll
            for (int i = 0; i < 8; ++i) ss = fmaf(x[s][i], x[s][i], ss);
        }
    }
    float scale = 1.f;
    if (p.q_normalise) {
        ss = warp_sum(ss);
        scale = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
    }
    float rs = 0.f;
#pragma unroll
    for (int s = 0; s < kBuildMaxSteps; ++s) {
        const int col = s * 256 + lane * 8;
        if (s < steps && col < p.d && col >= col_begin) {
            uint32_t packed[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __nv_bfloat16 lo = __float2bfloat16_rn(x[s][2 * i] * scale);
                const __nv_bfloat16 hi = __float2bfloat16_rn(x[s][2 * i + 1] * scale);
                const float flo = __bfloat162float(lo), fhi = __bfloat162float(hi);
                rs = fmaf(flo, flo, rs);
                rs = fmaf(fhi, fhi, rs);
                packed[i] = static_cast<uint32_t>(__bfloat16_as_ushort(lo)) |
                            (static_cast<uint32_t>(__bfloat16_as_ushort(hi)) << 16);
            }
            const int unit = (col - col_begin) >> 3, j = unit >> 3, c = unit & 7;
            *reinterpret_cast<uint4*>(q_tile + static_cast<size_t>(j) * slab_bytes + r * 128 + ((c ^ (r & 7)) << 4)) =
                make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
    }
    return warp_sum(rs);
}

// kCluster = 2 (shared-memory q-tile variant, even number of q-tiles): the two CTAs of a cluster work on the SAME bank
// split with DIFFERENT q-tiles; each loads half of every bank K-chunk and TMA-multicasts it into both CTAs' rings, which
// halves the L2 reads per MMA.  A ring slot is refilled only after BOTH consumers have released it (multicast
// tcgen05.commit, empty count = 2).  Tiles are statically split in this variant (both CTAs must walk the same tiles).
//
// kQTmem (D <= 512): the q-tile lives in TENSOR MEMORY (columns [0, D/2)) and is the MMA's TMEM A operand.  Only B then
// crosses the 128 B/cycle shared-memory port, and the shared memory the q-tile would occupy goes to the bank ring.  The
// accumulator ring is then 2 x 128 columns at [256, 512).
// kRegList (k + skip <= 8, the headline k = 5): the per-query list lives in REGISTERS (the best eight keys, sorted);
// only the pending buffer stays in shared memory.  Larger k keeps the list in shared memory as well.
// kFuseQ: see ScanParams::qsrc0 — replaces /root/reference/dataset/VQAFeatureDataset.py:189-191 in-kernel, for the
// tensor-memory q-tile (each epilogue thread converts its own row) and the shared-memory one (a warp per row, D <= 2048).
// kHybrid (with kQTmem): K-chunks 8.. of the q-tile are an SS-mode A operand in shared memory (brought in by TMA from
// prepared queries, or written by the warps from raw ones with kFuseQ; normalised raw queries are prepared first).  A
// compile-time switch because the MMA issue loop is issue-bound in the tensor-bound regime: a run-time branch per MMA
// group cost cfg4 (4096 x 1 M x 512) 15 %.
template <bool kDump, int kCluster, bool kQTmem, bool kFuseQ = false, bool kRegList = false, bool kHybrid = false>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_bank,
                 const ScanParams p, const TailParams tail) {
    extern __shared__ uint8_t smem_raw[];
    if (p.dbg_ring && threadIdx.x == 0)
        atomicMax(p.dbg_ring + 2 * (p.launch_seq & 63u), (1ull << 63) - ptx::globaltimer_ns());
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;   // SWIZZLE_128B atoms need 1024-byte alignment
    uint8_t* smem = smem_raw + (base - raw_addr);

    const ScanSmemLayout lay = scan_smem_layout(p.n_q_smem, p.q_box_rows, p.kk_pad, p.cand_cap, p.n_stages, p.sub_per_stage, p.n_epi_groups);
    const uint32_t q_smem = base + lay.q_off;
    const uint32_t stage_smem = base + lay.stage_off;
    uint64_t* lists = reinterpret_cast<uint64_t*>(smem + lay.list_off);
    float* bias_s = reinterpret_cast<float*>(smem + lay.bias_off);
    volatile uint64_t* tile_ring = reinterpret_cast<volatile uint64_t*>(smem + lay.ring_off);
    float* scratch = reinterpret_cast<float*>(smem + lay.scr_off);
    const uint32_t bar_base = base + lay.bar_off;
    const uint32_t bar_q = bar_base;
    auto bar_full = [&](int s) { return bar_base + 8u + 8u * s; };
    auto bar_empty = [&](int s) { return bar_base + 8u + 8u * (kMaxStages + s); };
    auto bar_tfull = [&](int b) { return bar_base + 8u + 8u * (2 * kMaxStages + b); };
    auto bar_tempty = [&](int b) { return bar_base + 8u + 8u * (2 * kMaxStages + kAccBufs + b); };
    const uint32_t bar_qs = bar_base + 8u + 8u * (2 * kMaxStages + 2 * kAccBufs);   // hybrid q-tile: the shared-memory half has landed (TMA)
    volatile uint32_t* tmem_slot =
        reinterpret_cast<volatile uint32_t*>(smem + lay.bar_off + (2 + 2 * kMaxStages + 2 * kAccBufs) * 8u);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int kBufs = kQTmem ? 2 : kAccBufs;               // accumulator ring depth
    constexpr uint32_t kAccCol0 = kQTmem ? 256u : 0u;          // first accumulator column
    constexpr bool kWarpsFillQ = kQTmem || kFuseQ;             // the epilogue warps (not TMA) bring the q-tile in

    // ---- which item is this CTA's
    const int item = blockIdx.x;
    const int split = item / p.n_qtiles;
    const int qt = item - split * p.n_qtiles;
    const int q0 = qt * p.q_tile;
    const int q_valid = min(p.q_tile, p.b_total - q0);
    const bool dynamic = kCluster == 1 && p.tile_ctr != nullptr;
    const int tile_begin = static_cast<int>(static_cast<long long>(split) * p.n_tiles / p.n_splits);
    const int tile_end = static_cast<int>(static_cast<long long>(split + 1) * p.n_tiles / p.n_splits);

    // ---- one-time setup
    if (threadIdx.x == 0) {
        ptx::mbar_init(bar_q, kWarpsFillQ ? 8 : 1);      // one arrive per epilogue warp, or the TMA producer's
        if constexpr (kHybrid) ptx::mbar_init(bar_qs, kFuseQ ? 8 : 1);   // the producer's TMA, or one arrive per filling warp
        for (int s = 0; s < p.n_stages; ++s) {
            ptx::mbar_init(bar_full(s), 1);
            ptx::mbar_init(bar_empty(s), kCluster);      // one release per consumer CTA of the cluster
        }
        for (int b = 0; b < kBufs; ++b) {
            ptx::mbar_init(bar_tfull(b), 1);
            ptx::mbar_init(bar_tempty(b), 4);   // one arrive per epilogue warp
        }

        ptx::fence_mbar_init();
        if constexpr (!kWarpsFillQ || kHybrid) ptx::prefetch_tensormap(&tmap_q);
        if (p.n_tiles > 0) ptx::prefetch_tensormap(&tmap_bank);
    }
    if (threadIdx.x < kTileRing) tile_ring[threadIdx.x] = 0ull;
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (kCluster > 1) ptx::cluster_sync();      // peers' barriers exist before anything remote touches them
    ptx::tc_fence_after();
    // Programmatic dependent launch (no-ops without the launch attribute): everything above touched only this CTA's
    // shared and tensor memory, so it may overlap the previous kernel of the stream — typically the previous retrieval
    // step, whose last CTAs are still in their tail when this one's first CTAs get an SM.  From here on global memory
    // is read (queries, control words the previous step re-zeroed, its result buffers are overwritten): wait for it.
    // The successor may be scheduled as soon as every CTA of this grid is resident (they all pass this point).
    ptx::griddep_wait();
    ptx::griddep_launch_dependents();
    const uint32_t tmem_base = *tmem_slot;
    // debug timeline (MPR_DEBUG_COUNTERS=1): dbg[8 + 16*cta + k] = globaltimer at event k of this CTA
    auto stamp = [&](int k) {
        if (p.dbg_ts) p.dbg_ts[24 * blockIdx.x + k] = ptx::globaltimer_ns_fenced();
    };
    if (threadIdx.x == 0) stamp(0);
    const uint32_t crank = kCluster > 1 ? ptx::cluster_ctarank() : 0u;
    constexpr uint16_t kClusterMask = static_cast<uint16_t>((1u << kCluster) - 1u);

    // Entry lt of the tile ring = ((lt + 1) << 32) | tile id, written by the producer when it starts loading this CTA's
    // lt-th tile (kTileEnd = no more tiles).  Consumers spin on the tag; the producer can run ahead of them by at most
    // the bank ring plus the accumulator ring (< kTileRing tiles), so an entry is never overwritten before it was read.
    auto ring_wait = [&](int lt) -> uint32_t {
        const uint64_t want = static_cast<uint64_t>(lt) + 1ull;
        uint64_t e = tile_ring[lt & (kTileRing - 1)];
        if ((e >> 32) != want) {
            const uint64_t t0 = ptx::globaltimer_ns();
            uint32_t polls = 0;
            while (((e = tile_ring[lt & (kTileRing - 1)]) >> 32) != want) {
                if ((++polls & 0x3FFu) == 0 && ptx::globaltimer_ns() - t0 > 4000000000ull) {
                    if (p.err) atomicCAS(p.err, 0, kErrRing);
                    __threadfence_system();
                    __trap();
                }
            }
        }
        return static_cast<uint32_t>(e);
    };

    // Producer and MMA warps run their loops with WARP-UNIFORM control flow (all 32 lanes wait on the barriers together)
    // and elect one lane only around the asynchronous instructions themselves.  Putting the whole loop under
    // `if (lane == 0)` makes every operand thread-variant for the compiler: it then wraps each UTCHMMA / UTMALDG in an
    // election loop with R2UR moves (~130 SASS instructions per K-chunk, measured: MMA issue-bound at 41 % tensor
    // activity with nothing waiting on data).
    if (warp == 0) {
        // =========================== TMA producer + tile scheduler ===========================
        if constexpr (!kWarpsFillQ) {
            if (ptx::elect_one()) {
                // resident q-tile: one 128-byte-wide slab per K-chunk
                const uint32_t slab_bytes = static_cast<uint32_t>(p.q_box_rows) * 128u;
                ptx::mbar_arrive_expect_tx(bar_q, slab_bytes * p.n_chunks);
                for (int j = 0; j < p.n_chunks; ++j)
                    ptx::tma_load_2d(q_smem + j * slab_bytes, &tmap_q, bar_q, j * kChunkK, q0, ptx::kEvictLast);
            }
            __syncwarp();
        }
        // hybrid q-tile: K-chunks 8.. come in by TMA into shared memory whose first 32 KiB serve as the scratch of the
        // warps' tensor-memory fill — they are fetched once that fill is through (bar_q), right before the first bank
        // chunk that needs them is requested
        bool q_smem_pending = kHybrid && !kFuseQ;      // raw queries: the warps write the shared-memory half themselves
        // streamed bank: a ring stage holds up to sub_per_stage 64-wide K sub-chunks and costs ONE barrier round-trip
        const int spp = p.sub_per_stage;
        const uint32_t t_limit = dynamic ? static_cast<uint32_t>(p.n_tiles) : static_cast<uint32_t>(tile_end);
        uint32_t t_cur;
        if (dynamic) {
            uint32_t g = 0;
            if (lane == 0) g = atomicAdd(p.tile_ctr + qt, 1u);
            t_cur = __shfl_sync(kFullMask, g, 0);
        } else {
            t_cur = static_cast<uint32_t>(tile_begin);
        }
        int s = 0;
        uint32_t ph = 0;
        for (int lt = 0;; ++lt) {
            const bool have = t_cur < t_limit;
            uint32_t t_next = t_cur + 1u;
            if (dynamic && have && lane == 0) t_next = atomicAdd(p.tile_ctr + qt, 1u);   // consumed after the loads below
            if (lane == 0) {
                if (have) {
                    tile_ring[lt & (kTileRing - 1)] = ((static_cast<uint64_t>(lt) + 1ull) << 32) | t_cur;
                } else {
                    for (int g = 0; g < kEpiGroups; ++g)      // every epilogue group's next tile index reads "end"
                        tile_ring[(lt + g) & (kTileRing - 1)] = ((static_cast<uint64_t>(lt + g) + 1ull) << 32) | kTileEnd;
                }
            }
            __syncwarp();
            if (!have) {
                // the MMA warp learns about the end through the ring as well, but it is parked on this barrier
                ptx::mbar_wait(bar_empty(s), ph ^ 1u, p.err, kErrEmpty);
                if (ptx::elect_one()) ptx::mbar_arrive(bar_full(s));
                __syncwarp();
                if (lane == 0) stamp(2);
                break;
            }
            const int t = static_cast<int>(t_cur);
            for (int j0 = 0; j0 < p.n_chunks; j0 += spp) {
                const int ns = min(spp, p.n_chunks - j0);
                if constexpr (kHybrid) {
                    if (q_smem_pending && j0 + ns > kQTmemChunks) {
                        ptx::mbar_wait(bar_q, 0, p.err, kErrQFull);          // every warp is through with the scratch
                        if (ptx::elect_one()) {
                            const uint32_t slab_bytes = static_cast<uint32_t>(p.q_box_rows) * 128u;
                            ptx::mbar_arrive_expect_tx(bar_qs, slab_bytes * p.n_q_smem);
                            for (int j = 0; j < p.n_q_smem; ++j)
                                ptx::tma_load_2d(q_smem + j * slab_bytes, &tmap_q, bar_qs, (kQTmemChunks + j) * kChunkK, q0, ptx::kEvictLast);
                        }
                        __syncwarp();
                        q_smem_pending = false;
                    }
                }
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
            t_cur = dynamic ? __shfl_sync(kFullMask, t_next, 0) : t_next;
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
        if (lane == 0) stamp(1);
        const int spp = p.sub_per_stage;
        int s = 0;
        uint32_t ph = 0;
        for (int lt = 0;; ++lt) {
            ptx::mbar_wait(bar_full(s), ph, p.err, kErrFull);      // first stage of tile lt, or the end marker
            if (lt == 1 && lane == 0) stamp(10);
            if (ring_wait(lt) == kTileEnd) break;
            const int buf = lt & (kBufs - 1);
            const uint32_t bph = (lt / kBufs) & 1u;
            ptx::mbar_wait(bar_tempty(buf), bph ^ 1u, p.err, kErrTmemEmpty);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + kAccCol0 + buf * kTileRows;
            for (int j0 = 0; j0 < p.n_chunks; j0 += spp) {
                const int ns = min(spp, p.n_chunks - j0);
                if (j0 > 0) ptx::mbar_wait(bar_full(s), ph, p.err, kErrFull);
                if constexpr (kHybrid) {
                    // a stage never straddles chunk 8 (sub_per_stage divides 8): the first stage beyond it needs the
                    // shared-memory half of the q-tile
                    if (lt == 0 && j0 == kQTmemChunks) ptx::mbar_wait(bar_qs, 0, p.err, kErrQFull);
                }
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
                                    if (!kHybrid || j0 < kQTmemChunks) {
                                        // 16 bf16 of K = 8 TMEM columns; sub-chunk j starts at column j*32
                                        ptx::umma_bf16_ts(d_tmem, tmem_base + j * (kChunkK / 2) + k * 8, db, idesc,
                                                          (j | k) != 0 ? 1u : 0u);
                                    } else {       // hybrid q-tile: this stage's A operand is in shared memory
                                        const uint64_t da = desc_hi | static_cast<uint64_t>(a_lo0 + (j - kQTmemChunks) * slab_lo + 2u * k);
                                        ptx::umma_bf16_ss(d_tmem, da, db, idesc, 1u);
                                    }
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
        // admission threshold in a register, appends the rare scores that beat it to its own pending buffer in shared
        // memory, and — when some lane's buffer runs full — every lane folds its own pending candidates into its own
        // sorted list.  No cross-lane traffic except one vote per 8 scores.
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
        // (with q_coop the list memory first serves as the q-tile fill's scratch; the lists are zeroed after it)
        if (!kRegList && active_group && !(kQTmem && p.q_coop)) for (int i = 0; i < kk; ++i) my_list[i] = 0ull;
        // thr = max(private threshold: score of this list's kk-th best, shared threshold: just below the minimum over
        // the query's ns published slot maxima).  `>` is exact for both: rows arrive in ascending order, so a later row
        // displaces the kk-th best only with a STRICTLY higher score; and ns distinct rows are known to score at least
        // the shared value, so only scores >= it (i.e. > its predecessor) can still matter, whatever their row.
        float thr_own = -CUDART_INF_F, thr_shared = -CUDART_INF_F;
        float thr = valid ? -CUDART_INF_F : CUDART_INF_F;
        float best = -CUDART_INF_F, best_published = -CUDART_INF_F;
        int n_pend = 0;
        const int flush_at = p.cand_cap - 8;           // the next group of 8 scores must always fit
        // The list is kept UNSORTED while the scan runs: an admitted candidate overwrites the current minimum and the new
        // minimum is found by one pass of independent loads (no store chain) — a sorted insert costs a dependent
        // load/compare/store per shifted element, which for k+s = 16..32 was most of the kernel's time (round 1: 43 % /
        // 24 % of the HBM roofline).  It is never sorted here: the tail's pool merge (tail.cuh) takes unordered lists.
        // register list (kRegList): the query's best EIGHT keys, sorted descending, 0 = empty.  Keeping eight whatever
        // k + skip <= 8 is makes every index static; the admission threshold is then the 8th best score (a superset of
        // the top-(k+skip) is kept, the first k+skip are written out).  An insert is position = number of entries above
        // the key (eight independent compares) followed by eight independent selects: ~90 instructions, depth ~10.
        uint64_t L[8];
        if constexpr (kRegList) {
#pragma unroll
            for (int j = 0; j < 8; ++j) L[j] = 0ull;
        }
        auto reg_insert = [&](uint64_t key) {      // caller: key > L[7]
            int pos = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) pos += L[j] > key ? 1 : 0;
#pragma unroll
            for (int j = 7; j >= 1; --j) L[j] = j > pos ? L[j - 1] : (j == pos ? key : L[j]);
            L[0] = pos == 0 ? key : L[0];
        };
        int n_list = 0, min_pos = 0;
        uint64_t min_key = 0ull;

        auto find_min = [&]() {
            uint64_t mk = my_list[0];
            int mp = 0;
#pragma unroll 8
            for (int j = 1; j < kk; ++j) {
                const uint64_t v = my_list[j];
                if (v < mk) { mk = v; mp = j; }
            }
            min_key = mk;
            min_pos = mp;
        };
        auto flush = [&]() {
            if (p.dbg) {
                atomicAdd(p.dbg + 0, static_cast<unsigned long long>(n_pend));
                if (lane == 0) atomicAdd(p.dbg + 1, 1ull);
            }
            if constexpr (kRegList) {
                // each lane folds ITS pending candidates: the loop runs max-over-lanes(n_pend) times, not once per bank row
                // in which some lane had a candidate (inserting straight from the score loop costs an insert per such row —
                // nearly every row of the first tiles, ~25 us of ramp-up per launch)
                for (int c = 0; c < n_pend; ++c) {
                    const uint2 cand = my_pend[c];
                    const uint64_t key = make_key(__uint_as_float(cand.x), cand.y);
                    if (key > L[7]) {
                        reg_insert(key);
                        if (p.dbg) atomicAdd(p.dbg + 3, 1ull);
                    }
                }
                n_pend = 0;
                thr_own = L[7] == 0ull ? -CUDART_INF_F : key_score(L[7]);
                if (valid) thr = fmaxf(thr_own, thr_shared);
                return;
            }
            for (int c = 0; c < n_pend; ++c) {
                const uint2 cand = my_pend[c];
                const uint64_t key = make_key(__uint_as_float(cand.x), cand.y);
                if (n_list < kk) {
                    my_list[n_list++] = key;
                    if (n_list == kk) find_min();
                } else if (key > min_key) {
                    my_list[min_pos] = key;
                    find_min();
                    if (p.dbg) atomicAdd(p.dbg + 3, 1ull);
                }
            }
            n_pend = 0;
            thr_own = n_list == kk ? key_score(min_key) : -CUDART_INF_F;
            if (valid) thr = fmaxf(thr_own, thr_shared);
        };
        if constexpr (kQTmem) {
            // This thread's query row -> its TMEM lane, columns [0, D/2): 8 bf16 (one uint4) fill 4 columns.  The two
            // epilogue groups split the row's 64-element chunks between them (both can reach every lane quadrant).
            const uint32_t q_taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
            const size_t qrow_idx = static_cast<size_t>(valid ? q0 + row : 0);
            if (p.q_coop) {
                // Warp-cooperative fill.  A thread pulling its own row reads 16 bytes per request from 32 different lines
                // (half of every sector fetched is wasted, and 148 CTAs pull the same 256 KB through L2 at once); here the
                // 32 rows x 64 elements of a chunk are read fully coalesced (8 lanes per row), rounded to bf16 and
                // transposed through 4 KiB of swizzled scratch so that each thread ends up with its own row's 128 bytes.
                // The sum of squares is taken by the owning thread over the rounded values in the same order as the
                // per-thread path, so q_bias_out is bit-identical.
                // scratch: the list memory, which is not in use yet — or, beside a hybrid q-tile, the head of the q-tile's
                // shared-memory half, which the producer fills only after this fill (bar_q)
                uint8_t* scr = smem + (kHybrid ? lay.q_off : lay.list_off) + (warp - 2) * 4096;
                const void* s0 = kFuseQ ? p.qsrc0 : static_cast<const void*>(p.q);
                const void* s1 = kFuseQ ? p.qsrc1 : nullptr;
                const int sd0 = kFuseQ ? p.qd0 : p.d, sd1 = kFuseQ ? p.qd1 : 0;
                const int sdt = kFuseQ ? p.q_dtype : static_cast<int>(kSrcBF16);
                float rs = 0.f;
                // all 8 (16 for fp32) 128-bit loads of a chunk are issued before the first conversion (kWide is a
                // compile-time constant so that the load loop has no branch the scheduler would not hoist loads across),
                // and the NEXT chunk's loads go out as soon as this chunk's registers are free — they are in flight while
                // this chunk's rows are read back, squared and stored to tensor memory
                uint4 raw[8][2];
                bool ok[8];
                auto load_chunk = [&](auto wide_tag, int c) {
                    constexpr bool kWide = decltype(wide_tag)::value;        // fp32 source: two 128-bit loads per unit
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int unit_idx = it * 32 + lane;
                        const int rl = unit_idx >> 3, u = unit_idx & 7;
                        ok[it] = quad * 32 + rl < q_valid;
                        const size_t grow = static_cast<size_t>(q0 + (ok[it] ? quad * 32 + rl : 0));
                        const int col = c * kChunkK + u * 8;
                        const bool first = col < sd0;
                        const size_t off = first ? grow * sd0 + col : grow * sd1 + (col - sd0);
                        const uint4* src = reinterpret_cast<const uint4*>(
                            static_cast<const uint8_t*>(first ? s0 : s1) + off * (kWide ? 4 : 2));
                        raw[it][0] = __ldg(src);
                        if constexpr (kWide) raw[it][1] = __ldg(src + 1);
                    }
                };
                auto store_chunk = [&](auto wide_tag) {
                    constexpr bool kWide = decltype(wide_tag)::value;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int unit_idx = it * 32 + lane;
                        const int rl = unit_idx >> 3, u = unit_idx & 7;
                        uint32_t pk[4];
                        if constexpr (kWide) {
                            const uint32_t f[8] = {raw[it][0].x, raw[it][0].y, raw[it][0].z, raw[it][0].w,
                                                   raw[it][1].x, raw[it][1].y, raw[it][1].z, raw[it][1].w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const __nv_bfloat16 lo = __float2bfloat16_rn(__uint_as_float(f[2 * i]));
                                const __nv_bfloat16 hi = __float2bfloat16_rn(__uint_as_float(f[2 * i + 1]));
                                pk[i] = static_cast<uint32_t>(__bfloat16_as_ushort(lo)) |
                                        (static_cast<uint32_t>(__bfloat16_as_ushort(hi)) << 16);
                            }
                        } else {
                            const uint32_t h[4] = {raw[it][0].x, raw[it][0].y, raw[it][0].z, raw[it][0].w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                if (sdt == kSrcF16) {
                                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
                                    pk[i] = static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(f.x))) |
                                            (static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(f.y))) << 16);
                                } else {
                                    pk[i] = h[i];                            // bf16 rows pass through
                                }
                            }
                        }
                        const uint4 out = ok[it] ? make_uint4(pk[0], pk[1], pk[2], pk[3]) : make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(scr + rl * 128 + ((u ^ (rl & 7)) << 4)) = out;
                    }
                };
                const int n_tmem_chunks = min(p.n_chunks, kQTmemChunks);
                const bool wide = sdt == kSrcF32;
                if (grp < n_tmem_chunks) {
                    if (wide) load_chunk(std::true_type{}, grp);
                    else      load_chunk(std::false_type{}, grp);
                }
                for (int c = grp; c < n_tmem_chunks; c += kEpiGroups) {
                    if (wide) store_chunk(std::true_type{});
                    else      store_chunk(std::false_type{});
                    __syncwarp();
                    if (c + kEpiGroups < n_tmem_chunks) {
                        if (wide) load_chunk(std::true_type{}, c + kEpiGroups);
                        else      load_chunk(std::false_type{}, c + kEpiGroups);
                    }
                    uint32_t w[32];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint4 t = *reinterpret_cast<const uint4*>(scr + lane * 128 + ((u ^ (lane & 7)) << 4));
                        w[4 * u + 0] = t.x; w[4 * u + 1] = t.y; w[4 * u + 2] = t.z; w[4 * u + 3] = t.w;
                    }
                    if constexpr (kFuseQ) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float flo = __uint_as_float(w[i] << 16), fhi = __uint_as_float(w[i] & 0xFFFF0000u);
                            rs = fmaf(flo, flo, rs);
                            rs = fmaf(fhi, fhi, rs);
                        }
                    }
                    ptx::tmem_st_32x32b_x32(q_taddr + c * (kChunkK / 2), w);
                    __syncwarp();      // the scratch is rewritten by the next chunk
                }
                if constexpr (kFuseQ && kHybrid) {
                    // raw queries beside a hybrid q-tile: once every warp is through with the scratch (it lives in the
                    // shared-memory half), the eight warps write that half themselves, a warp per row (kernel 1's row
                    // routine restricted to the columns beyond 512), and release it to the MMA warp
                    ptx::named_bar_sync(5, 256);
                    const uint32_t slab_bytes = static_cast<uint32_t>(p.q_box_rows) * 128u;
                    for (int r = warp - 2; r < p.q_box_rows; r += 8) {
                        const bool real = q0 + r < p.b_total && r < q_valid;
                        const float rs2 = prep_query_row_to_smem(p, q0 + r, real, smem + lay.q_off, slab_bytes, r, lane,
                                                                 kQTmemChunks * kChunkK);
                        if (lane == 0 && r < kUmmaM) scratch[2 * kUmmaM + r] = real ? rs2 : 0.f;
                    }
                    ptx::fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core's async proxy
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(bar_qs);
                }
                if constexpr (kFuseQ) {
                    if (split == 0 && p.q_bias_out) {                   // -0.5*|q|^2 for return_dists: the two halves meet here
                        scratch[grp * kUmmaM + row] = rs;
                        ptx::named_bar_sync(3, 256);
                        if (grp == 0 && valid)
                            p.q_bias_out[q0 + row] = -0.5f * (rs + scratch[kUmmaM + row] + (kHybrid ? scratch[2 * kUmmaM + row] : 0.f));
                    }
                }
                if constexpr (!kRegList) {
                    ptx::named_bar_sync(4, 256);       // every warp is through with its scratch: the memory becomes lists
                    if (active_group) for (int i = 0; i < kk; ++i) my_list[i] = 0ull;
                }
                if constexpr (kHybrid && !kFuseQ) ptx::fence_proxy_async_smem();   // scratch accesses before the TMA that overwrites it
            } else if constexpr (kFuseQ) {
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
                for (int c0 = grp * 32; c0 < p.d / 2; c0 += 64) {
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
                if (split == 0 && p.q_bias_out) {                   // -0.5*|q|^2 for return_dists: the two halves meet here
                    scratch[grp * kUmmaM + row] = rs;
                    ptx::named_bar_sync(3, 256);
                    if (grp == 0 && valid) p.q_bias_out[q0 + row] = -0.5f * (rs + scratch[kUmmaM + row]);
                }
            } else {
                const uint4* qrow = reinterpret_cast<const uint4*>(p.q + qrow_idx * p.d);
                // two 64-element chunks per round: 16 independent 128-bit loads in flight before the first store
                for (int c0 = grp * 32; c0 < p.d / 2; c0 += 128) {
                    uint32_t w0[32], w1[32];
                    const bool second = c0 + 64 < p.d / 2;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        uint4 t = make_uint4(0u, 0u, 0u, 0u), t2 = make_uint4(0u, 0u, 0u, 0u);
                        if (valid) t = __ldg(qrow + c0 / 4 + u);
                        if (valid && second) t2 = __ldg(qrow + (c0 + 64) / 4 + u);
                        w0[4 * u + 0] = t.x; w0[4 * u + 1] = t.y; w0[4 * u + 2] = t.z; w0[4 * u + 3] = t.w;
                        w1[4 * u + 0] = t2.x; w1[4 * u + 1] = t2.y; w1[4 * u + 2] = t2.z; w1[4 * u + 3] = t2.w;
                    }
                    ptx::tmem_st_32x32b_x32(q_taddr + c0, w0);
                    if (second) ptx::tmem_st_32x32b_x32(q_taddr + c0 + 64, w1);
                }
            }
            ptx::tmem_wait_st();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_q);
        } else if constexpr (kFuseQ) {
            // shared-memory q-tile filled by the eight epilogue warps, a warp per row (coalesced 1 KiB reads)
            const uint32_t slab_bytes = static_cast<uint32_t>(p.q_box_rows) * 128u;
            for (int r = warp - 2; r < p.q_box_rows; r += 8) {
                const bool real = q0 + r < p.b_total;
                const float rs = prep_query_row_to_smem(p, q0 + r, real, smem + lay.q_off, slab_bytes, r, lane);
                if (lane == 0 && real && r < q_valid && split == 0 && p.q_bias_out) p.q_bias_out[q0 + r] = -0.5f * rs;
            }
            ptx::fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_q);
        }

        auto load_bias = [&](uint32_t t) -> float {
            const uint32_t r = t * kTileRows + ep_tid;
            return (r < static_cast<uint32_t>(p.n_local)) ? __ldg(p.bias + r) : -CUDART_INF_F;   // -inf masks rows past the shard end
        };
        // shared threshold: minimum over this query's ns slots (0 = some slot still empty -> no bound yet)
        // slot of this list: group 1 is offset by ns/2, so that the group-0 lists ALONE feed every slot — group 0 owns a
        // CTA's first tile, and its first-tile wait below must not depend on second tiles (one tile time later)
        const int my_slot_ = p.gthr ? (((split + grp * (p.ns >> 1)) % p.ns) << p.thr_rep_log2) +
                                          ((split / p.ns) & ((1 << p.thr_rep_log2) - 1)) : 0;      // word index: slot * R + replica
        const uint32_t* my_gthr = p.gthr ? p.gthr + (valid ? q0 + row : q0) : nullptr;     // slot i at my_gthr[i * b_total]
        const size_t thr_pitch = static_cast<size_t>(p.b_total);
        auto apply_shared = [&](uint32_t m) {
            // strictly-below-m in the ordered domain: scores >= m stay admissible (their row may still win a tie)
            if (m > 0x00800000u) {
                thr_shared = fmaxf(thr_shared, __uint_as_float(ordered_to_f32(m - 1u)));
                if (valid) thr = fmaxf(thr_own, thr_shared);
                if (p.dbg && valid) atomicAdd(p.dbg + 5, 1ull);
            }
        };
        const int thr_words = p.ns << p.thr_rep_log2;
        constexpr int kThrN = kRegList ? 16 : 32;       // words held in registers (the host keeps ns << r within it)
        // kThrN words (static indices; word i = slot (i >> r), replica (i & (R-1))) -> minimum over the slots of the maximum
        // over each slot's replicas.  Unused words must be 0xFFFFFFFF... for the min and are never the max of a slot.
        auto fold_words = [&](uint32_t (&v)[kThrN]) -> uint32_t {
            if (p.thr_rep_log2 >= 1) {
#pragma unroll
                for (int i = 0; i < kThrN; i += 2) v[i] = max(v[i], v[i + 1]);
            }
            if (p.thr_rep_log2 >= 2) {
#pragma unroll
                for (int i = 0; i < kThrN; i += 4) v[i] = max(v[i], v[i + 2]);
            }
            const int step = 1 << p.thr_rep_log2;
            uint32_t m = 0xFFFFFFFFu;
#pragma unroll
            for (int i = 0; i < kThrN; ++i)
                if ((i & (step - 1)) == 0 && i < thr_words) m = min(m, v[i]);
            return m;
        };
        auto load_words = [&](uint32_t (&v)[kThrN]) {
#pragma unroll
            for (int i = 0; i < kThrN; ++i) v[i] = i < thr_words ? __ldcg(my_gthr + i * thr_pitch) : 0u;
        };
        auto slots_min = [&]() -> uint32_t {        // one loaded-L2 round trip (1-2 us under a full scan)
            uint32_t v[kThrN];
            load_words(v);
            return fold_words(v);
        };
        auto refresh_shared = [&]() { apply_shared(slots_min()); };      // blocking form
        // split form: the loads are issued when a tile starts and consumed when it ends, so their latency hides under the
        // tile's own work (a blocking refresh per tile cost ~2 us each during the ramp-up, ~15 us per launch)
        uint32_t thr_pre[kThrN];
        bool thr_pending = false;
        auto refresh_issue = [&]() {
            load_words(thr_pre);
            thr_pending = true;
        };
        auto refresh_consume = [&]() {
            apply_shared(fold_words(thr_pre));
            thr_pending = false;
        };
        // publish this list's best score: the minimum over a query's slots bounds its global kk-th best from below
        auto publish_best = [&]() {
            if (valid && best > best_published) {
                atomicMax(p.gthr + static_cast<size_t>(my_slot_) * p.b_total + q0 + row, f32_to_ordered(__float_as_uint(best)));
                best_published = best;
            }
        };
        // With n_epi_groups == 1 group 1 idles in the main loop: it only helped to load the q-tile.
        const int n_groups = p.n_epi_groups;
        bool have_bias = false;
        float next_bias = 0.f;
        int done_tiles = 0, next_refresh = 1;

        if (active_group) {
            for (int lt = grp;; lt += n_groups) {
                const uint32_t t = ring_wait(lt);
                if (t == kTileEnd) break;
                const int buf = lt & (kBufs - 1);
                const uint32_t bph = (lt / kBufs) & 1u;
                if (!have_bias) next_bias = load_bias(t);
                // bias tiles are double-buffered per group: a fast warp may stage tile lt+2 while a slow one still reads lt
                float* bias_tile = bias_s + (grp * 2 + ((lt / n_groups) & 1)) * kTileRows;
                bias_tile[ep_tid] = next_bias;
                ptx::named_bar_sync(1 + grp, 128);
                if (done_tiles == 1 && grp == 0 && lane == 0 && quad == 0) stamp(21);
                {   // the producer is normally a tile or two ahead: fetch the next tile's bias under this tile's work
                    const uint64_t e = tile_ring[(lt + n_groups) & (kTileRing - 1)];
                    have_bias = (e >> 32) == static_cast<uint64_t>(lt + n_groups) + 1ull && static_cast<uint32_t>(e) != kTileEnd;
                    if (have_bias) next_bias = load_bias(static_cast<uint32_t>(e));
                }
                if (my_gthr && done_tiles >= next_refresh) {
                    refresh_issue();
                    next_refresh = done_tiles + max(1, done_tiles >> 1);
                }

                ptx::mbar_wait(bar_tfull(buf), bph, p.err, kErrTmemFull);
                ptx::tc_fence_after();
                if (done_tiles == 0 && grp == 0 && lane == 0 && quad == 0) stamp(15);

                if (warp_has_work) {
                    const uint32_t row_base = t * kTileRows;
                    const uint32_t acc_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + kAccCol0 + buf * kTileRows;
                    const bool legacy_first = p.first_wait_ns < 0;
                    bool first_bound = false;      // the wait below ended with every slot filled
                    if (done_tiles == 0 && my_gthr && !legacy_first) {
                        // First tile of a list: nothing is known about the query yet, every CTA starts blind at the same
                        // moment, and a blind row costs an insert.  So the tile is walked TWICE (the accumulators stay in
                        // tensor memory): pass 1 only takes the maximum of its 128 scores and publishes it; then the warp
                        // waits — briefly, bounded — until every threshold slot of its queries holds some list's maximum
                        // (a slot is fed by ~n_lists/ns lists, the earliest one suffices), which bounds the query's kk-th
                        // best from below by the minimum of ns maxima over >= 128 rows each; pass 2 is the ordinary walk
                        // and admits a handful of rows instead of all of them.
                        float cm4[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll 1
                        for (int c0 = 0; c0 < kTileRows; c0 += 32) {
                            uint32_t v[32];
                            ptx::tmem_ld_32x32b_x32(acc_addr + c0, v);
                            ptx::tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {      // four independent max chains
                                const float4 bb = *reinterpret_cast<const float4*>(bias_tile + c0 + i);
                                cm4[0] = fmaxf(cm4[0], __uint_as_float(v[i + 0]) + bb.x);
                                cm4[1] = fmaxf(cm4[1], __uint_as_float(v[i + 1]) + bb.y);
                                cm4[2] = fmaxf(cm4[2], __uint_as_float(v[i + 2]) + bb.z);
                                cm4[3] = fmaxf(cm4[3], __uint_as_float(v[i + 3]) + bb.w);
                            }
                        }
                        const float cm = fmaxf(fmaxf(cm4[0], cm4[1]), fmaxf(cm4[2], cm4[3]));
                        best = fmaxf(best, cm);
                        publish_best();
                        if (grp == 0 && lane == 0 && quad == 0) stamp(23);
                        const uint64_t w0 = ptx::globaltimer_ns();
                        for (;;) {
                            const uint32_t m = slots_min();
                            const bool filled = !valid || m != 0u;
                            if (__all_sync(kFullMask, filled)) {
                                apply_shared(m);
                                first_bound = true;
                                break;
                            }
                            if (ptx::globaltimer_ns() - w0 >= static_cast<uint64_t>(p.first_wait_ns)) break;
                        }
                        if (grp == 0 && lane == 0 && quad == 0) stamp(16);
                    }
#pragma unroll 1
                    for (int c0 = 0; c0 < kTileRows; c0 += 32) {
                        uint32_t v[32];
                        ptx::tmem_ld_32x32b_x32(acc_addr + c0, v);
                        ptx::tmem_wait_ld();
                        if (done_tiles == 0 && my_gthr && legacy_first) {
                            // legacy start: make this chunk's best score public BEFORE working on it
                            float cm = -CUDART_INF_F;
#pragma unroll
                            for (int i = 0; i < 32; ++i) cm = fmaxf(cm, __uint_as_float(v[i]) + bias_tile[c0 + i]);
                            best = fmaxf(best, cm);
                            publish_best();
                        }
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
                                if (p.dbg && lane == 0) atomicAdd(p.dbg + 2, 1ull);
                                const uint32_t gidx = p.idx_base + row_base + c0 + g * 8;
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    if (s[i] > thr) {
                                        my_pend[n_pend] = make_uint2(__float_as_uint(s[i]), gidx + i);
                                        ++n_pend;
                                    }
                                }
                                if (m > thr) best = fmaxf(best, m);
                                if (__any_sync(kFullMask, n_pend > flush_at)) flush();
                            }
                        }
                        // legacy start, and the fallback when the wait above ran out: one blocking refresh after the first
                        // 32 rows (shared-memory lists, whose blind inserts cost hundreds of cycles each: after every 32)
                        if (done_tiles == 0 && my_gthr && !first_bound && (!kRegList || c0 == 0)) refresh_shared();
                        if (done_tiles == 0 && grp == 0 && lane == 0 && quad == 0) stamp(17 + min(c0 >> 5, 2));
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar_tempty(buf));
                if (done_tiles == 0 && grp == 0 && lane == 0 && quad == 0) stamp(20);
                if (thr_pending) refresh_consume();
                if (my_gthr) publish_best();
                ++done_tiles;
                if (p.dbg && lane == 0 && warp_has_work) atomicAdd(p.dbg + 4, 1ull);
                if (done_tiles == 1 && lane == 0 && quad == 0) stamp(11 + grp);
                if (done_tiles == 4 && lane == 0 && quad == 0) stamp(13 + grp);
            }
        }

        if (lane == 0 && quad == 0) stamp(3 + 2 * grp);
        // partial result of this (split, group, q-tile): part_keys[q0 + row][split * 2 + grp][rank]
        if (warp_has_work) {
            // list-major: every thread's kk keys are contiguous (one or a few full sectors per thread; rank-major cost a
            // separate 32-byte sector per 8-byte key — 40 MB of sector writes per launch at k + skip = 32)
            const size_t n_lists = static_cast<size_t>(p.n_splits) * kEpiGroups;
            uint64_t* dst = p.part_keys + (static_cast<size_t>(q0 + row) * n_lists + split * kEpiGroups + grp) * kk;
            if constexpr (kRegList) {
                if (active_group) flush();
                if (valid) {                      // the register list is already sorted: its first k + skip entries go out
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (i < kk) dst[i] = active_group ? L[i] : 0ull;
                }
            } else {
                if (active_group) flush();
                if (lane == 0 && quad == 0 && grp == 0) stamp(22);
                if (valid)                       // as it lies (unordered, empty slots are 0): the tail's pool merge needs no order
                    for (int i = 0; i < kk; ++i) dst[i] = active_group ? my_list[i] : 0ull;
            }
        }
        if (lane == 0 && quad == 0) stamp(4 + 2 * grp);
    }

    // ---- teardown
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (kCluster > 1) ptx::cluster_sync();      // no CTA leaves while a peer may still signal its barriers
    if (threadIdx.x == 0) stamp(7);
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }

    // ---- fused tail: once every CTA's partial lists are in memory, warps finish whole queries (merge -> [exchange] ->
    //      vote -> prompt ids); the last CTA out re-zeroes the control block.  Requires a co-resident grid.
    if constexpr (!kDump && kCluster == 1) {
        if (p.fused_tail) {
            if (threadIdx.x == 0) {
                __threadfence();
                atomicAdd(tail.ctrl, 1u);
                if (ld_acquire_gpu_u32(tail.ctrl) < gridDim.x) {
                    const uint64_t t0 = ptx::globaltimer_ns();
                    uint32_t polls = 0;
                    while (ld_acquire_gpu_u32(tail.ctrl) < gridDim.x) {
                        if ((++polls & 0x3FFu) == 0 && ptx::globaltimer_ns() - t0 > 8000000000ull) {
                            if (p.err) atomicCAS(p.err, 0, kErrGridBarrier);
                            __threadfence_system();
                            __trap();
                        }
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) stamp(8);
            uint32_t e = 0;
            if (tail.xchg.world > 1) e = *reinterpret_cast<volatile uint32_t*>(tail.xchg.peers.buf[tail.xchg.rank]) + 1u;
            // the list memory of the scan is free now: 10 x 32 keys of it carry the warps' partial merges
            uint64_t* sbuf = reinterpret_cast<uint64_t*>(smem + lay.list_off);
            for (int q = blockIdx.x; q < tail.b; q += gridDim.x) block_tail_query<kScanWarps>(tail, q, e, sbuf, warp, lane);
            __syncthreads();
            if (threadIdx.x == 0) stamp(9);
            if (threadIdx.x == 0) tail_ticket(tail, e, gridDim.x);
        }
    }
    if (p.dbg_ring && threadIdx.x == 0) atomicMax(p.dbg_ring + 2 * (p.launch_seq & 63u) + 1, ptx::globaltimer_ns());
}

}  // namespace mpr
