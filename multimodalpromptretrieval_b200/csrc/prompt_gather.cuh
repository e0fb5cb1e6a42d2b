// Kernel 3 — retrieved rows -> answer vote -> quantifier bucket -> prompt token ids (one warp per query), stand-alone
// form for callers that already hold an index matrix.  Inside a retrieval step the same warp function
// (warp_vote_and_gather, tail.cuh) runs in the scan kernel's tail.
//
// Device-side restatement of the tail of retrieve_closest_qa_pairs and of the tokeniser call in prepare_input:
//   answers[i][j] = retrieval_answers[idx[i, j]]                     /root/reference/dataset/VQAFeatureDataset.py:199
//   majority vote; ties -> the answer whose first occurrence has the lowest rank                       :216-222
//   certainty = max_count / n_votes; bucket = buckets[int(certainty * 5)]                              :223-226
//   "I believe the answer is {bucket} {answer}" | "The most frequent answer is {answer}"               :228,230
//   sentence = "Answer the {task} question: " + question + retrieved  (no separating space)
//   tokenizer(padding="longest", max_length, truncation=True)        /root/reference/architectures/T5VisionModel.py:153-167
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "tail.cuh"

namespace mpr {

__global__ void __launch_bounds__(128) prompt_gather_kernel(const PromptParams p) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= p.b) return;
    const int k = p.kk - p.skip;
    const PromptPrefetch f = prompt_prefetch(p, q, lane);
    const int row = lane < k ? p.idx[q * p.kk + p.skip + lane] : -1;
    warp_vote_and_gather(p, q, row, lane, f);
}

}  // namespace mpr
