// The tail of a retrieval step — everything that follows the bank scan for ONE query, written as warp-level device
// functions so that it can run (a) inside the scan kernel itself after its grid barrier (one launch per step) and
// (b) as the stand-alone kernels of the multi-launch path (merge_topk.cuh, prompt_gather.cuh, tail_kernel below):
//
//   1. k-way merge of the per-(split, epilogue group) partial lists the scan left in its workspace
//   2. multi-GPU only: push the local top-kk into every peer's exchange buffer over NVLink (plain P2P stores + a
//      release flag per query), wait for every rank's delivery of the same query, merge the `world` lists
//   3. answers of the retrieved rows -> majority vote -> quantifier bucket -> prompt token ids
//
// Replaces the second half of torch.argsort(...)[:, s:s+k] (/root/reference/dataset/VQAFeatureDataset.py:195,197), the
// answer gather / vote / quantifier sentence (:199,215-230) and the tokenizer call of
// /root/reference/architectures/T5VisionModel.py:153-167.  The cross-GPU exchange is new (the reference is
// single-device, /root/reference/main.py:58-61).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <math_constants.h>

#include "ptx.cuh"
#include "topk_key.cuh"

namespace mpr {

// ------------------------------------------------------------------------------------------------ k-way merge
// Candidate (list l, rank i) lives at base[l*stride_l + i*stride_i]; every list is sorted descending over i, 0 = empty.
// The walk is rank-major (all lists' best, then all second-best, ...).  Because every list is sorted, a rank at which no
// list contributes ends the merge: no later rank can beat the threshold either.  Returns the merged list, one element per
// lane (lanes >= kk are scratch).  kCoherent: bypass L1 (the lists were written by other SMs / GPUs during this launch).
constexpr int kMergeUnroll = 8;

template <bool kCoherent>
__device__ __forceinline__ uint64_t load_key(const uint64_t* p) {
    if constexpr (kCoherent) return __ldcg(reinterpret_cast<const unsigned long long*>(p));
    else return *p;
}

template <bool kCoherent>
__device__ __forceinline__ uint64_t warp_merge_lists(const uint64_t* __restrict__ base, int n_lists, long long stride_l,
                                                     long long stride_i, int kk, int lane) {
    uint64_t elem = 0ull;        // lane i holds element i of the running list
    uint64_t kth = 0ull;         // admission threshold: max(element kk-1, floor)
    uint64_t floor = 0ull;       // a key known to be below the final kk-th best (0 = none)
    for (int i = 0; i < kk; ++i) {
        bool admitted = false;
        const uint64_t* rank_base = base + static_cast<long long>(i) * stride_i;
        for (int l0 = 0; l0 < n_lists; l0 += 32 * kMergeUnroll) {
            uint64_t key[kMergeUnroll];
#pragma unroll
            for (int u = 0; u < kMergeUnroll; ++u) {          // all loads of the chunk in flight before any use
                const int l = l0 + u * 32 + lane;
                key[u] = l < n_lists ? load_key<kCoherent>(rank_base + static_cast<long long>(l) * stride_l) : 0ull;
            }
            if (i == 0 && l0 == 0 && n_lists >= 64) {
                // Pivot: the kk-th largest of the 32 lanes' local maxima is a key that at least kk candidates reach, so
                // everything below it can skip the serial insert loop (with ~300 lists that is all but ~2*kk of them).
                uint64_t lm = key[0];
#pragma unroll
                for (int u = 1; u < kMergeUnroll; ++u) lm = key[u] > lm ? key[u] : lm;
                int rank = 0;
                for (int o = 1; o < 32; ++o) rank += shfl_u64(lm, (lane + o) & 31) > lm ? 1 : 0;
                const unsigned who = __ballot_sync(kFullMask, rank == kk - 1 && lm != 0ull);
                if (who) {
                    floor = shfl_u64(lm, __ffs(who) - 1) - 1ull;
                    kth = floor;
                }
            }
#pragma unroll
            for (int u = 0; u < kMergeUnroll; ++u) {
                if (l0 + u * 32 >= n_lists) break;
                unsigned pending = __ballot_sync(kFullMask, key[u] > kth);
                while (pending) {
                    const int src = __ffs(pending) - 1;
                    pending &= pending - 1;
                    const uint64_t cand = shfl_u64(key[u], src);
                    if (cand > kth) {    // uniform: the threshold may have moved since the ballot
                        elem = warp_list_insert(elem, cand, lane);
                        const uint64_t last = shfl_u64(elem, kk - 1);
                        kth = last > floor ? last : floor;
                        admitted = true;
                    }
                }
            }
        }
        if (!admitted) break;
    }
    return elem;
}

// Top-kk of an UNORDERED pool of m keys (0 = empty), contiguous at `pool`: what the scan's (split, group) lists of one
// query are.  Two streaming passes: (1) every lane's maximum -> the kk-th largest of the 32 lane maxima is a key that at
// least kk keys reach, so everything below it is out; (2) the few survivors are inserted into the warp's sorted list.
// Neither pass depends on the order inside the per-CTA lists, so the scan writes them unsorted, and unlike a rank-by-rank
// walk over sorted lists there is no chain of dependent L2 round trips (14 us per query at k + skip = 32).
// The scan's shared thresholds end as ns >= kk maxima over disjoint row sets: kk distinct rows score at least their
// minimum, so a key whose score is strictly below it cannot be among the query's top kk (ties stay in).  `word` = this
// lane's threshold word (word = slot * R + replica; lanes >= ns * R pass anything); returns the floor key or 0.
__device__ __forceinline__ uint64_t floor_from_threshold_words(uint32_t word, int ns, int rep_log2, int lane) {
    const int words = ns << rep_log2;
    uint32_t mn = lane < words ? word : 0u;
    if (rep_log2 >= 1) mn = max(mn, __shfl_xor_sync(kFullMask, mn, 1));      // maximum over a slot's replicas
    if (rep_log2 >= 2) mn = max(mn, __shfl_xor_sync(kFullMask, mn, 2));
    if (lane >= words) mn = 0xFFFFFFFFu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(kFullMask, mn, o));
    return mn > 0x00800000u ? (static_cast<uint64_t>(mn) << 32) - 1ull : 0ull;
}

// `thr_word`: where this lane's threshold word lives (nullptr = no floor from the thresholds); it is loaded AFTER the
// pool's first round of loads has been issued, so both ride the same L2 round trip.  With a floor the pivot pass is
// skipped for kk > 8, where the kk-th largest of 32 lane maxima is a weak pivot anyway.
// `floor0`: a key known to lie below the pool's kk-th best (0 = none), e.g. from the scan's shared thresholds; with one
// the pivot pass is skipped for kk > 8, where the kk-th largest of 32 lane maxima is a weak pivot anyway.
template <bool kCoherent>
__device__ __forceinline__ uint64_t warp_merge_pool(const uint64_t* __restrict__ pool, int m, int kk, int lane,
                                                    const uint32_t* thr_word = nullptr, int ns = 0, int rep_log2 = 0) {
    uint64_t key0[kMergeUnroll];      // the first round of loads goes out before anything else
#pragma unroll
    for (int u = 0; u < kMergeUnroll; ++u) {
        const int i = u * 32 + lane;
        key0[u] = i < m ? load_key<kCoherent>(pool + i) : 0ull;
    }
    uint64_t floor0 = 0ull;
    if (thr_word) floor0 = floor_from_threshold_words(lane < (ns << rep_log2) ? __ldcg(thr_word) : 0u, ns, rep_log2, lane);
    uint64_t floor = floor0;
    const bool want_pivot = floor0 == 0ull || kk <= 8;
    auto pivot_of = [&](uint64_t lm) -> uint64_t {      // the kk-th largest of the 32 lane maxima, minus one
        int rank = 0;
        for (int o = 1; o < 32; ++o) rank += shfl_u64(lm, (lane + o) & 31) > lm ? 1 : 0;
        const unsigned who = __ballot_sync(kFullMask, rank == kk - 1 && lm != 0ull);
        return who ? shfl_u64(lm, __ffs(who) - 1) - 1ull : 0ull;
    };
    auto insert_batch = [&](const uint64_t (&key)[kMergeUnroll], uint64_t& elem, uint64_t& kth) {
#pragma unroll
        for (int u = 0; u < kMergeUnroll; ++u) {
            unsigned pending = __ballot_sync(kFullMask, key[u] > kth);
            while (pending) {
                const int src = __ffs(pending) - 1;
                pending &= pending - 1;
                const uint64_t cand = shfl_u64(key[u], src);
                if (cand > kth) {    // uniform: the threshold may have moved since the ballot
                    elem = warp_list_insert(elem, cand, lane);
                    const uint64_t last = shfl_u64(elem, kk - 1);
                    kth = last > floor ? last : floor;
                }
            }
        }
    };
    if (m <= 32 * kMergeUnroll) {
        // the whole pool fits the warp's registers: ONE round of loads serves the pivot and the inserts (the headline
        // shape, 296 lists x 5 keys over ten warps, is 160 keys per warp)
        if (want_pivot) {
            uint64_t lm = key0[0];
#pragma unroll
            for (int u = 1; u < kMergeUnroll; ++u) lm = key0[u] > lm ? key0[u] : lm;
            const uint64_t pivot = pivot_of(lm);
            floor = pivot > floor ? pivot : floor;
        }
        uint64_t elem = 0ull, kth = floor;
        insert_batch(key0, elem, kth);
        return elem;
    }
    if (want_pivot) {
        uint64_t lm = 0ull;
#pragma unroll
        for (int u = 0; u < kMergeUnroll; ++u) lm = key0[u] > lm ? key0[u] : lm;
        for (int i0 = 32 * kMergeUnroll; i0 < m; i0 += 32 * kMergeUnroll) {
            uint64_t key[kMergeUnroll];
#pragma unroll
            for (int u = 0; u < kMergeUnroll; ++u) {
                const int i = i0 + u * 32 + lane;
                key[u] = i < m ? load_key<kCoherent>(pool + i) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < kMergeUnroll; ++u) lm = key[u] > lm ? key[u] : lm;
        }
        const uint64_t pivot = pivot_of(lm);
        floor = pivot > floor ? pivot : floor;
    }
    uint64_t elem = 0ull, kth = floor;
    insert_batch(key0, elem, kth);
    for (int i0 = 32 * kMergeUnroll; i0 < m; i0 += 32 * kMergeUnroll) {
        uint64_t key[kMergeUnroll];
#pragma unroll
        for (int u = 0; u < kMergeUnroll; ++u) {
            const int i = i0 + u * 32 + lane;
            key[u] = i < m ? load_key<kCoherent>(pool + i) : 0ull;
        }
        insert_batch(key, elem, kth);
    }
    return elem;
}

__device__ __forceinline__ void store_merged(uint64_t elem, int q, int kk, int lane, uint64_t* out_keys, float* out_score,
                                             int32_t* out_idx) {
    if (lane < kk) {
        const size_t o = static_cast<size_t>(q) * kk + lane;
        if (out_keys) out_keys[o] = elem;
        if (out_score) out_score[o] = elem == 0ull ? -CUDART_INF_F : key_score(elem);
        if (out_idx) out_idx[o] = key_row(elem);
    }
}

// ------------------------------------------------------------------------------------------------ peer-memory exchange
// Every rank owns one exchange buffer of identical layout, mapped into every peer (symmetric memory), zero-filled once:
//
//   byte 0      u32 epoch                  last exchange this rank has fully consumed
//   byte 4      u32 done                   tail tickets of the launch in flight (stand-alone tail kernel only)
//   byte 1024   u64 word[4][world][cap][2] candidate lists ([q][kk] keys inside a [rank] block, two words per key),
//                                          four buffers used in turn (exchange e uses buffer e & 3)
//
// Exchange e = epoch + 1.  A key travels as TWO self-validating 8-byte words, {low half | e << 32} and
// {high half | e << 32}: an aligned 8-byte store is single-copy atomic, so a reader that finds tag e in a word has that
// word's payload — no fence, no separate flag, no second NVLink round trip (the flag + fence.sys protocol this replaces
// cost a store round trip, a system fence and a flag round trip per step).  For query q a warp stores its kk keys into
// word[e&3][my_rank] of EVERY rank's buffer (one 16-byte store per key and peer), then polls word[e&3][0..world) of its
// OWN buffer until every word carries tag e and merges the world lists (keys are unique — they embed the global row — so
// the merge is order-independent and ties still resolve to the lower row).  The last warp of the launch publishes
// epoch = e.  Stale words carry an older tag (or 0) whatever batch shape wrote them.
// Why four buffers are enough.  In-kernel collection: a rank can start exchange e+1 while a slow peer still reads the
// words of e, but it cannot reach e+2 before that peer has delivered its own e+1 words, which it does only after its
// merge of e (two buffers would do).  DEFERRED collection (the step kernel only pushes; `xchg_finish_kernel` on a side
// stream collects, so that the NVLink latency hides under the next step's scan): the host lets step j+2 be queued only
// after the finish of step j has run on this rank, so a rank reaching the push of e+4 has finished e+2, hence holds
// every peer's words of e+2, hence every peer has launched e+2, hence finished e — nobody still reads buffer e & 3.
// The waiting warp only ever waits on OTHER GPUs, never on a kernel that must be co-scheduled on its own.
// Sharded search is therefore a COLLECTIVE call: every rank must issue the same sequence of searches.
constexpr int kXchgMaxWorld = 16;
constexpr int kXchgDataOff = 1024;
constexpr int kXchgBuffers = 4;
constexpr int kXchgMaxPerLane = kXchgMaxWorld * 32 / 32;      // world * kk <= 512 keys, spread over 32 lanes
constexpr int kErrXchgTimeout = 201;

struct XchgPeers {
    unsigned char* buf[kXchgMaxWorld];
};

struct XchgParams {
    int world, rank, cap;          // world <= 1: no exchange
    unsigned long long timeout_ns; // how long a warp waits for a peer before it gives up (status = kErrXchgTimeout)
    int mode;                      // tuning bits: 1 = system fence after the stores, 2 = poll with L2 (.cg) loads
    XchgPeers peers;               // peers.buf[rank] = this rank's own buffer
};

__host__ __device__ inline size_t xchg_bytes(int world, int cap) {
    return static_cast<size_t>(kXchgDataOff) + static_cast<size_t>(kXchgBuffers) * world * cap * 2ull * sizeof(uint64_t);
}

__device__ __forceinline__ void st_relaxed_sys_v2_u64(uint64_t* p, uint64_t a, uint64_t b) {
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_relaxed_sys_v2_u64(const uint64_t* p, uint64_t& a, uint64_t& b) {
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Top-kk of up to 32 * kN keys held in registers (0 = empty), any order: pivot from the lane maxima, then inserts.
template <int kN>
__device__ __forceinline__ uint64_t warp_merge_regs(const uint64_t (&key)[kN], int kk, int lane) {
    uint64_t lm = key[0];
#pragma unroll
    for (int u = 1; u < kN; ++u) lm = key[u] > lm ? key[u] : lm;
    int rank = 0;
    for (int o = 1; o < 32; ++o) rank += shfl_u64(lm, (lane + o) & 31) > lm ? 1 : 0;
    const unsigned who = __ballot_sync(kFullMask, rank == kk - 1 && lm != 0ull);
    const uint64_t floor = who ? shfl_u64(lm, __ffs(who) - 1) - 1ull : 0ull;
    uint64_t elem = 0ull, kth = floor;
#pragma unroll
    for (int u = 0; u < kN; ++u) {
        unsigned pending = __ballot_sync(kFullMask, key[u] > kth);
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            const uint64_t cand = shfl_u64(key[u], src);
            if (cand > kth) {    // uniform: the threshold may have moved since the ballot
                elem = warp_list_insert(elem, cand, lane);
                const uint64_t last = shfl_u64(elem, kk - 1);
                kth = last > floor ? last : floor;
            }
        }
    }
    return elem;
}

// Push of one query's merged local list (`elem`, one element per lane) into every rank's buffer (own included).
__device__ __forceinline__ void warp_exchange_push(const XchgParams& x, uint32_t e, uint64_t elem, int q, int kk, int lane) {
    const uint64_t tag = static_cast<uint64_t>(e) << 32;
    const size_t buf_off = static_cast<size_t>(e & (kXchgBuffers - 1)) * x.world * x.cap * 2;      // in words
    if (lane < kk) {
        const uint64_t w0 = (elem & 0xFFFFFFFFull) | tag, w1 = (elem >> 32) | tag;
        const size_t off = buf_off + (static_cast<size_t>(x.rank) * x.cap + static_cast<size_t>(q) * kk + lane) * 2;
        for (int r = 0; r < x.world; ++r)
            st_relaxed_sys_v2_u64(reinterpret_cast<uint64_t*>(x.peers.buf[r] + kXchgDataOff) + off, w0, w1);
    }
    if (x.mode & 1) __threadfence_system();
}

// Collection of one query's lists from all ranks out of this rank's own buffer; returns the global list.  On a timeout
// *status receives kErrXchgTimeout, *timed_out is set and whatever has arrived (the own list at least) is merged.
__device__ __forceinline__ uint64_t warp_exchange_collect(const XchgParams& x, uint32_t e, int q, int kk, int lane,
                                                          int* status, bool* timed_out) {
    const size_t buf_off = static_cast<size_t>(e & (kXchgBuffers - 1)) * x.world * x.cap * 2;      // in words
    const uint64_t* mine = reinterpret_cast<const uint64_t*>(x.peers.buf[x.rank] + kXchgDataOff) + buf_off;
    const int total = x.world * kk;
    uint64_t ent[kXchgMaxPerLane];
    unsigned need = 0u;
#pragma unroll
    for (int j = 0; j < kXchgMaxPerLane; ++j) {
        ent[j] = 0ull;
        if (lane + 32 * j < total) need |= 1u << j;
    }
    const uint64_t t0 = ptx::globaltimer_ns();
    bool ok = true;
    *timed_out = false;
    for (uint32_t polls = 0;; ++polls) {
#pragma unroll
        for (int j = 0; j < kXchgMaxPerLane; ++j) {
            if (need & (1u << j)) {
                const int idx = lane + 32 * j, r = idx / kk, i = idx - r * kk;
                uint64_t a, b;
                const uint64_t* src = mine + (static_cast<size_t>(r) * x.cap + static_cast<size_t>(q) * kk + i) * 2;
                if (x.mode & 2) {
                    const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(src));
                    a = v.x; b = v.y;
                } else {
                    ld_relaxed_sys_v2_u64(src, a, b);
                }
                if ((a >> 32) == e && (b >> 32) == e) {
                    ent[j] = (b << 32) | (a & 0xFFFFFFFFull);
                    need &= ~(1u << j);
                }
            }
        }
        if (!__any_sync(kFullMask, need != 0u)) break;
        if ((polls & 0x3Fu) == 0x3Fu && ptx::globaltimer_ns() - t0 > x.timeout_ns) ok = false;
        if (!__all_sync(kFullMask, ok)) {
            if (lane == 0 && status) atomicMax(status, kErrXchgTimeout);
            *timed_out = true;
            break;
        }
    }
    return warp_merge_regs<kXchgMaxPerLane>(ent, kk, lane);
}

// Exchange of one query's merged local list with all ranks inside one kernel; returns the global list, or — when a
// peer did not deliver in time (status = kErrXchgTimeout) — the LOCAL list.
__device__ __forceinline__ uint64_t warp_exchange(const XchgParams& x, uint32_t e, uint64_t elem, int q, int kk, int lane,
                                                  int* status) {
    warp_exchange_push(x, e, elem, q, kk, lane);
    bool timed_out;
    const uint64_t merged = warp_exchange_collect(x, e, q, kk, lane, status, &timed_out);
    return timed_out ? elem : merged;
}

// ------------------------------------------------------------------------------------------------ vote + prompt ids
// Tokenisation is done by concatenating pre-tokenised segments (sentencepiece never merges across whitespace):
//   prefix_q  = tokens("Answer the {task} question: " + question + "I" | "The")      host, per query
//   seg 0 / 1 = tokens("believe the answer is") / tokens("most frequent answer is")
//   seg 2..7  = tokens(bucket b),   seg 8+a = tokens(answer a)                        host, once per bank
constexpr int kSegQuant = 0, kSegPlain = 1, kSegBucket0 = 2, kSegAnswer0 = 8;

struct PromptParams {
    const int32_t* idx;         // [b][kk] global bank rows, -1 = none (stand-alone kernel only)
    int b, kk, skip;            // k = kk - skip votes per query, taken from ranks skip..kk-1
    const int32_t* answer_id;   // [n_total] interned answer of every bank row; nullptr = no prompt stage
    const uint8_t* bucket_lut;  // [(k+1)*(k+1)]: lut[n_votes*(k+1) + max_count] = int(max_count/n_votes*5)
    const int32_t* prefix_ids;  // CSR over queries
    const int32_t* prefix_off;  // [b+1]
    const int32_t* seg_ids;     // CSR over segments
    const int32_t* seg_off;     // [8 + n_answers + 1]
    int use_quantifier;
    int pad_id, eos_id;
    int max_len;                // tokenizer max_length (truncation), eos included
    int out_stride;             // row pitch of input_ids / attention_mask
    long long* input_ids;       // [b][out_stride]
    long long* attention_mask;  // [b][out_stride]
    int32_t* out_len;           // [b] tokens incl. eos
    int32_t* maj_answer;        // [b] answer id of the vote winner (-1 if no votes)
    int32_t* maj_count;         // [b]
    int32_t* bucket;            // [b] 0..5
    int32_t* ret_answer;        // [b][k] answer ids in rank order (or nullptr)
};

// Everything the prompt stage reads that does NOT depend on the vote, fetched up front so that it is in flight while the
// lists are merged: the query's prefix range, the offsets of the constant / bucket segments (lane i: seg_off[i], i <= 8)
// and the bucket-table row for a full vote (lane c: lut[k*(k+1) + c]).  What remains behind the vote is the chain
// answer_id[row] -> seg_off[answer] -> token ids.
struct PromptPrefetch {
    int pre0, pre1, seg, lut_full;
};

__device__ __forceinline__ PromptPrefetch prompt_prefetch(const PromptParams& p, int q, int lane) {
    PromptPrefetch f;
    const int k = p.kk - p.skip;
    f.pre0 = p.prefix_off[q];
    f.pre1 = p.prefix_off[q + 1];
    f.seg = lane <= kSegAnswer0 ? p.seg_off[lane] : 0;
    f.lut_full = lane <= k ? p.bucket_lut[k * (k + 1) + lane] : 0;
    return f;
}

// `row` = bank row retrieved at rank skip + lane (lanes >= k: ignored), -1 = none.
__device__ __forceinline__ void warp_vote_and_gather(const PromptParams& p, int q, int row, int lane,
                                                     const PromptPrefetch& f) {
    const int k = p.kk - p.skip;
    // ---- gather the answers of the retrieved rows   (VQAFeatureDataset.py:199)
    int a = -1;
    if (lane < k) {
        if (row >= 0) a = __ldg(p.answer_id + row);
        if (p.ret_answer) p.ret_answer[q * k + lane] = a;
    }
    const unsigned voters = __ballot_sync(kFullMask, a >= 0);
    const int n_votes = __popc(voters);

    // ---- majority vote, ties to the earliest first occurrence   (:216-222)
    const unsigned same = __match_any_sync(kFullMask, a) & voters;
    int rank_key = -1;
    if (a >= 0) rank_key = __popc(same) * 64 + (63 - (__ffs(same) - 1));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rank_key = max(rank_key, __shfl_xor_sync(kFullMask, rank_key, o));
    int maj = -1, cnt = 0, bkt = 0;
    if (rank_key >= 0) {      // warp-uniform
        cnt = rank_key >> 6;
        const int first = 63 - (rank_key & 63);
        maj = __shfl_sync(kFullMask, a, first);
        // :223-226; the row for n_votes == k is already in registers (cnt <= k <= 31)
        bkt = (n_votes == k && cnt < 32) ? __shfl_sync(kFullMask, f.lut_full, cnt) : p.bucket_lut[n_votes * (k + 1) + cnt];
    }
    if (lane == 0) {
        p.maj_answer[q] = maj;
        p.maj_count[q] = cnt;
        p.bucket[q] = bkt;
    }

    // ---- token assembly: prefix | const | [bucket] | answer | </s> | pad...   (:228,230; T5VisionModel.py:153-167)
    const int pre0 = f.pre0, len_pre = f.pre1 - f.pre0;
    const int seg_c = p.use_quantifier ? kSegQuant : kSegPlain;
    const int c0 = __shfl_sync(kFullMask, f.seg, seg_c), len_c = __shfl_sync(kFullMask, f.seg, seg_c + 1) - c0;
    int b0 = 0, len_b = 0;
    {
        const int bs = __shfl_sync(kFullMask, f.seg, kSegBucket0 + bkt), be = __shfl_sync(kFullMask, f.seg, kSegBucket0 + bkt + 1);
        if (p.use_quantifier) { b0 = bs; len_b = be - bs; }
    }
    int a0 = 0, len_a = 0;
    if (maj >= 0) { a0 = p.seg_off[kSegAnswer0 + maj]; len_a = p.seg_off[kSegAnswer0 + maj + 1] - a0; }
    const int body = min(len_pre + len_c + len_b + len_a, p.max_len - 1);   // HF truncation keeps room for </s>
    if (lane == 0) p.out_len[q] = body + 1;

    long long* ids = p.input_ids + static_cast<size_t>(q) * p.out_stride;
    long long* msk = p.attention_mask + static_cast<size_t>(q) * p.out_stride;
    for (int pos = lane; pos < p.out_stride; pos += 32) {
        int tok = p.pad_id;
        if (pos < body) {
            int o = pos;
            if (o < len_pre) tok = p.prefix_ids[pre0 + o];
            else if ((o -= len_pre) < len_c) tok = p.seg_ids[c0 + o];
            else if ((o -= len_c) < len_b) tok = p.seg_ids[b0 + o];
            else tok = p.seg_ids[a0 + (o - len_b)];
        } else if (pos == body) {
            tok = p.eos_id;
        }
        ids[pos] = tok;
        msk[pos] = pos <= body ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------ one query's tail
// Scan workspace control block (first bytes of the caller's workspace; all-zero before the first launch, and every
// launch leaves it all-zero again):
//   u32 arrived     grid barrier of the fused tail / nothing in the multi-launch path
//   u32 done        tail completion tickets
//   u32 tile_ctr[n_qtiles]   dynamic tile scheduler, one counter per q-tile
//   u32 gthr[ns][b]          shared admission thresholds, slot-major (see scan_topk.cuh)
struct TailParams {
    int b, kk;
    const uint64_t* part_keys;   // [b][n_lists][kk], each list unordered
    int n_lists;
    uint32_t* ctrl;              // -> arrived, done
    uint32_t* tile_ctr;
    int n_tile_ctr;
    uint32_t* gthr;              // [ns * R][b]
    int ns;
    int thr_rep_log2;            // 2^r replica words per slot (see ScanParams)
    int use_floor;               // pool merge: keys below the scan's final shared threshold are dropped unseen
    uint64_t* out_keys;          // [b][kk] (each may be nullptr)
    float* out_score;
    int32_t* out_idx;
    int* status;                 // device word: 0 = ok (may be nullptr)
    int defer_xchg;              // world > 1: push only; xchg_finish_kernel (side stream) collects, votes and writes the outputs
    uint32_t* finish_epoch;      // deferred: the launch's last ticket leaves its exchange epoch here for the finish kernel
    XchgParams xchg;
    PromptParams prompt;
};

// Runs the whole tail for query q on one thread block of kWarps warps (every thread of the block must call it): the
// warps split the query's candidate pool between them (the pool merge is a stream of L2 reads; one warp keeps too few
// of them in flight), leave their sorted top-kk in shared memory (sbuf: kWarps x 32 keys), and warp 0 merges those and
// finishes the query.  `e` = exchange epoch of this launch (ignored when world <= 1).
template <int kWarps>
__device__ __forceinline__ void block_tail_query(const TailParams& t, int q, uint32_t e, uint64_t* sbuf, int warp, int lane) {
    const int m = t.n_lists * t.kk;
    const int per = ((m + kWarps - 1) / kWarps + 31) & ~31;
    const int lo = min(m, warp * per), hi = min(m, lo + per);
    const uint64_t* pool = t.part_keys + static_cast<size_t>(q) * m;
    PromptPrefetch pf = {0, 0, 0, 0};
    if (warp == 0 && t.prompt.answer_id && !(t.xchg.world > 1 && t.defer_xchg)) pf = prompt_prefetch(t.prompt, q, lane);
    const uint32_t* thr_word = (t.gthr && t.use_floor) ? t.gthr + static_cast<size_t>(min(lane, (t.ns << t.thr_rep_log2) - 1)) * t.b + q
                                                       : nullptr;
    const uint64_t mine = warp_merge_pool<true>(pool + lo, hi - lo, t.kk, lane, thr_word, t.ns, t.thr_rep_log2);
    sbuf[warp * 32 + lane] = lane < t.kk ? mine : 0ull;
    __syncthreads();
    if (warp == 0) {
        uint64_t elem = warp_merge_lists<false>(sbuf, kWarps, 32ll, 1ll, t.kk, lane);
        if (t.gthr && lane < (t.ns << t.thr_rep_log2)) t.gthr[static_cast<size_t>(lane) * t.b + q] = 0u;     // leave the thresholds zeroed
        if (t.xchg.world > 1 && t.defer_xchg) {
            warp_exchange_push(t.xchg, e, elem, q, t.kk, lane);      // the rest of the query happens in xchg_finish_kernel
        } else {
        if (t.xchg.world > 1) elem = warp_exchange(t.xchg, e, elem, q, t.kk, lane, t.status);
        store_merged(elem, q, t.kk, lane, t.out_keys, t.out_score, t.out_idx);
        if (t.prompt.answer_id) {
            const uint64_t src = shfl_u64(elem, min(lane + t.prompt.skip, 31));
            const int row = (lane < t.kk - t.prompt.skip) ? key_row(src) : -1;
            warp_vote_and_gather(t.prompt, q, row, lane, pf);
        }
        }
    }
    __syncthreads();
}

// Completion ticket: the last warp group to finish re-zeroes the control block and publishes the exchange epoch.
// Call with exactly `n_tickets` callers per launch (one thread each).
__device__ __forceinline__ void tail_ticket(const TailParams& t, uint32_t e, uint32_t n_tickets) {
    __threadfence();
    const uint32_t done = atomicAdd(t.ctrl + 1, 1u);
    if (done == n_tickets - 1u) {
        for (int i = 0; i < t.n_tile_ctr; ++i) t.tile_ctr[i] = 0u;
        t.ctrl[0] = 0u;
        t.ctrl[1] = 0u;
        if (t.xchg.world > 1) {
            if (t.defer_xchg && t.finish_epoch) *t.finish_epoch = e;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(t.xchg.peers.buf[t.xchg.rank]) = e;
        }
    }
}

// Stand-alone tail (multi-launch path: grids larger than one wave cannot hold a grid barrier): one block per query.
constexpr int kTailWarps = 4;

__global__ void __launch_bounds__(kTailWarps * 32) tail_kernel(const TailParams t) {
    __shared__ uint64_t sbuf[kTailWarps * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t e = 0;
    if (t.xchg.world > 1) e = *reinterpret_cast<volatile uint32_t*>(t.xchg.peers.buf[t.xchg.rank]) + 1u;
    for (int q = blockIdx.x; q < t.b; q += gridDim.x) block_tail_query<kTailWarps>(t, q, e, sbuf, warp, lane);
    if (threadIdx.x == 0) tail_ticket(t, e, gridDim.x);
}

// Deferred half of a sharded step (TailParams::defer_xchg): one warp per query collects the ranks' lists from this
// rank's exchange buffer — by the time it runs, on a side stream next to the NEXT step's scan, they have normally all
// arrived — merges them, writes the search outputs, votes and gathers the prompt ids.
// One warp per block and few registers: it has to fit NEXT TO a resident step CTA (10 warps x 168 registers leave
// 256 registers free on two of an SM's four register-file partitions and 5.6 K on the other two; a 4-warp block, one
// warp per partition, never fits and would only run once the next step's kernel has left).
constexpr int kFinishWarps = 1;

__global__ void __maxnreg__(64) xchg_finish_kernel(const TailParams t) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= t.b) return;
    const uint32_t e = *reinterpret_cast<volatile const uint32_t*>(t.finish_epoch);
    PromptPrefetch pf = {0, 0, 0, 0};
    if (t.prompt.answer_id) pf = prompt_prefetch(t.prompt, q, lane);
    bool timed_out;
    const uint64_t elem = warp_exchange_collect(t.xchg, e, q, t.kk, lane, t.status, &timed_out);
    store_merged(elem, q, t.kk, lane, t.out_keys, t.out_score, t.out_idx);
    if (t.prompt.answer_id) {
        const uint64_t src = shfl_u64(elem, min(lane + t.prompt.skip, 31));
        const int row = (lane < t.kk - t.prompt.skip) ? key_row(src) : -1;
        warp_vote_and_gather(t.prompt, q, row, lane, pf);
    }
}

}  // namespace mpr
