// Kernel 3 — retrieved rows -> answer vote -> quantifier bucket -> prompt token ids (one warp per query).
//
// Device-side restatement of the tail of retrieve_closest_qa_pairs and of the tokeniser call in prepare_input:
//   answers[i][j] = retrieval_answers[idx[i, j]]                     /root/reference/dataset/VQAFeatureDataset.py:199
//   majority vote; ties -> the answer whose first occurrence has the lowest rank                       :216-222
//   certainty = max_count / n_votes; bucket = buckets[int(certainty * 5)]                              :223-226
//   "I believe the answer is {bucket} {answer}" | "The most frequent answer is {answer}"               :228,230
//   sentence = "Answer the {task} question: " + question + retrieved  (no separating space)
//   tokenizer(padding="longest", max_length, truncation=True)        /root/reference/architectures/T5VisionModel.py:153-167
//
// Tokenisation is done by concatenating pre-tokenised segments (sentencepiece never merges across whitespace):
//   prefix_q  = tokens("Answer the {task} question: " + question + "I" | "The")      host, per query
//   seg 0 / 1 = tokens("believe the answer is") / tokens("most frequent answer is")
//   seg 2..7  = tokens(bucket b),   seg 8+a = tokens(answer a)                        host, once per bank
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mpr {

constexpr int kSegQuant = 0, kSegPlain = 1, kSegBucket0 = 2, kSegAnswer0 = 8;

struct PromptParams {
    const int32_t* idx;         // [b][kk] global bank rows, -1 = none
    int b, kk, skip;            // k = kk - skip votes per query, taken from ranks skip..kk-1
    const int32_t* answer_id;   // [n_total] interned answer of every bank row
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

__global__ void __launch_bounds__(128) prompt_gather_kernel(const PromptParams p) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= p.b) return;
    const int k = p.kk - p.skip;

    // ---- gather the answers of the retrieved rows (lane j = rank j after the skip)
    int a = -1;
    if (lane < k) {
        const int r = p.idx[q * p.kk + p.skip + lane];
        if (r >= 0) a = __ldg(p.answer_id + r);
        if (p.ret_answer) p.ret_answer[q * k + lane] = a;
    }
    const unsigned voters = __ballot_sync(0xFFFFFFFFu, a >= 0);
    const int n_votes = __popc(voters);

    // ---- majority vote, ties to the earliest first occurrence
    const unsigned same = __match_any_sync(0xFFFFFFFFu, a) & voters;
    int rank_key = -1;
    if (a >= 0) rank_key = __popc(same) * 64 + (63 - (__ffs(same) - 1));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rank_key = max(rank_key, __shfl_xor_sync(0xFFFFFFFFu, rank_key, o));
    int maj = -1, cnt = 0, bkt = 0;
    if (rank_key >= 0) {
        cnt = rank_key >> 6;
        const int first = 63 - (rank_key & 63);
        maj = __shfl_sync(0xFFFFFFFFu, a, first);
        bkt = p.bucket_lut[n_votes * (k + 1) + cnt];
    }
    if (lane == 0) {
        p.maj_answer[q] = maj;
        p.maj_count[q] = cnt;
        p.bucket[q] = bkt;
    }

    // ---- token assembly: prefix | const | [bucket] | answer | </s> | pad...
    const int pre0 = p.prefix_off[q], len_pre = p.prefix_off[q + 1] - pre0;
    const int seg_c = p.use_quantifier ? kSegQuant : kSegPlain;
    const int c0 = p.seg_off[seg_c], len_c = p.seg_off[seg_c + 1] - c0;
    int b0 = 0, len_b = 0;
    if (p.use_quantifier) { b0 = p.seg_off[kSegBucket0 + bkt]; len_b = p.seg_off[kSegBucket0 + bkt + 1] - b0; }
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

}  // namespace mpr
