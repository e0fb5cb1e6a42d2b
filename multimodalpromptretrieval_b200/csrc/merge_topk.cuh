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

// in_keys: n_lists lists; list l of query q starts at in_keys + l*list_stride + q*kk  (kk keys, 0 = empty).
// out_keys [b][kk] (may be nullptr), out_score [b][kk] (may be nullptr), out_idx [b][kk] (may be nullptr).
__global__ void __launch_bounds__(128)
merge_topk_kernel(const uint64_t* __restrict__ in_keys, int n_lists, long long list_stride, int b, int kk,
                  uint64_t* __restrict__ out_keys, float* __restrict__ out_score, int32_t* __restrict__ out_idx) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= b) return;

    uint64_t elem = 0ull;        // lane i holds element i of the running list
    uint64_t kth = 0ull;         // element kk-1 (the admission threshold)
    const int total = n_lists * kk;
    for (int c0 = 0; c0 < total; c0 += 32) {
        const int c = c0 + lane;
        uint64_t key = 0ull;
        if (c < total) {
            const int l = c / kk;
            key = in_keys[static_cast<long long>(l) * list_stride + static_cast<long long>(q) * kk + (c - l * kk)];
        }
        unsigned pending = __ballot_sync(kFullMask, key > kth);
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            const uint64_t cand = shfl_u64(key, src);
            if (cand > kth) {    // uniform: the threshold may have moved since the ballot
                elem = warp_list_insert(elem, cand, lane);
                kth = shfl_u64(elem, kk - 1);
            }
        }
    }
    if (lane < kk) {
        const size_t o = static_cast<size_t>(q) * kk + lane;
        if (out_keys) out_keys[o] = elem;
        if (out_score) out_score[o] = elem == 0ull ? -CUDART_INF_F : key_score(elem);
        if (out_idx) out_idx[o] = key_row(elem);
    }
}

}  // namespace mpr
