// Kernel 4 — k-way merge of sorted candidate lists (one warp per query), stand-alone form.
//
// Used after an NCCL all-gather of every rank's [B][kk] candidates (RetrievalBank(exchange="nccl")) to fold the
// per-rank lists ([world][B][kk]) into the global top-k.  Keys order by (score desc, global row asc), so the result is
// identical for any split / rank count.  The merge of the scan's own per-split partial lists and the peer-memory
// exchange run inside the scan kernel's tail (tail.cuh); the warp-level merge itself is shared (warp_merge_lists).
//
// This is the cross-GPU half of torch.argsort(...)[:, s:s+k] (/root/reference/dataset/VQAFeatureDataset.py:195,197).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "tail.cuh"

namespace mpr {

// Candidate (list l, query q, rank i) lives at in_keys[l*stride_l + q*stride_q + i*stride_i]; every list is sorted
// descending over i, 0 = empty.  out_keys / out_score / out_idx [b][kk] (each may be nullptr).
__global__ void __launch_bounds__(128)
merge_topk_kernel(const uint64_t* __restrict__ in_keys, int n_lists, long long stride_l, long long stride_q,
                  long long stride_i, int b, int kk, uint64_t* __restrict__ out_keys, float* __restrict__ out_score,
                  int32_t* __restrict__ out_idx) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= b) return;
    const uint64_t elem = warp_merge_lists<false>(in_keys + static_cast<long long>(q) * stride_q, n_lists, stride_l,
                                                  stride_i, kk, lane);
    store_merged(elem, q, kk, lane, out_keys, out_score, out_idx);
}

}  // namespace mpr
