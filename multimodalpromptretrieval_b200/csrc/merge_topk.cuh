// Kernel 4 — k-way merge of sorted candidate lists (one warp per query).
//
// Used twice: (a) to fold the per-split partial lists the scan kernel leaves in its workspace, and (b) after the
// NCCL allgather, to fold the per-rank lists ([world][B][kk]) into the global top-k.  Keys order by
// (score desc, global row asc), so the result is identical for any split / rank count.
//
// This is the cross-CTA / cross-GPU half of torch.argsort(...)[:, s:s+k]
// (/root/reference/dataset/VQAFeatureDataset.py:195,197).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <math_constants.h>

#include "topk_key.cuh"

namespace mpr {

// Candidate (list l, query q, rank i) lives at in_keys[l*stride_l + q*stride_q + i*stride_i]; every list is sorted
// descending over i, 0 = empty.  Two layouts are in use:
//   scan partials  [q][i][split]  (stride_l = 1, stride_i = n_lists, stride_q = kk*n_lists): a rank-major walk is one
//                                  contiguous, fully coalesced stream per query;
//   rank lists     [rank][q][i]   (stride_l = b*kk, stride_q = kk, stride_i = 1): what the all-gather delivers.
// The walk is rank-major (all lists' best, then all second-best, ...).  Because every list is sorted, a rank at which
// no list contributes ends the merge: no later rank can beat the threshold either.
// out_keys [b][kk] (may be nullptr), out_score [b][kk] (may be nullptr), out_idx [b][kk] (may be nullptr).
constexpr int kMergeUnroll = 8;

__global__ void __launch_bounds__(128)
merge_topk_kernel(const uint64_t* __restrict__ in_keys, int n_lists, long long stride_l, long long stride_q,
                  long long stride_i, int b, int kk, uint64_t* __restrict__ out_keys, float* __restrict__ out_score,
                  int32_t* __restrict__ out_idx) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= b) return;

    uint64_t elem = 0ull;        // lane i holds element i of the running list
    uint64_t kth = 0ull;         // element kk-1 (the admission threshold)
    const uint64_t* base = in_keys + static_cast<long long>(q) * stride_q;
    for (int i = 0; i < kk; ++i) {
        bool admitted = false;
        const uint64_t* rank_base = base + static_cast<long long>(i) * stride_i;
        for (int l0 = 0; l0 < n_lists; l0 += 32 * kMergeUnroll) {
            uint64_t key[kMergeUnroll];
#pragma unroll
            for (int u = 0; u < kMergeUnroll; ++u) {          // all loads of the chunk in flight before any use
                const int l = l0 + u * 32 + lane;
                key[u] = l < n_lists ? rank_base[static_cast<long long>(l) * stride_l] : 0ull;
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
                        kth = shfl_u64(elem, kk - 1);
                        admitted = true;
                    }
                }
            }
        }
        if (!admitted) break;
    }
    if (lane < kk) {
        const size_t o = static_cast<size_t>(q) * kk + lane;
        if (out_keys) out_keys[o] = elem;
        if (out_score) out_score[o] = elem == 0ull ? -CUDART_INF_F : key_score(elem);
        if (out_idx) out_idx[o] = key_row(elem);
    }
}

}  // namespace mpr
