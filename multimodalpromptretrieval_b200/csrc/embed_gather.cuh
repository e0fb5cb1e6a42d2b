// Kernel 5 — prompt token ids -> T5 input embeddings, with the image tokens prepended and the mask built in the same pass
// (SURVEY.md §8f N4).  Device-side restatement of the end of prepare_input:
//   question_embedding = self.T5_model.shared(encoding["input_ids"])          /root/reference/architectures/T5VisionModel.py:169
//   image_attn_mask    = torch.ones((B, n_image_tokens))                                                             :172
//   attention_mask     = torch.cat((image_attn_mask, encoding.attention_mask), axis=1)   (float32 by promotion)      :173
//   combined_embedding = torch.cat((image_embeddings, question_embedding), axis=1)                                   :176
// The ids come straight from kernel 3 / the scan tail on the device, so the tokenizer's H2D copy of the ids and the two
// cat kernels disappear.  Forward only: the gradient of `shared` (trainable in T5VisionModelFrozen.py:24) is an
// index_add over the same ids and stays on stock PyTorch (host wrapper: embed.py).
//
// One warp per output row; a row is `row_vec` 16-byte vectors (hidden 512 fp32 = 128 vectors, 4 per lane) — pure
// HBM/L2-bound copy: bytes = B*(n_image+L)*hidden*esize read + the same written.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mpr {

constexpr int kErrTokenRange = 301;

struct EmbedParams {
    const long long* input_ids;       // [b][in_stride], first `len` columns used
    const long long* attention_mask;  // [b][in_stride]
    int b, len, in_stride;
    const uint4* table;               // [vocab][row_vec]
    int vocab, row_vec;
    const uint4* image_tokens;        // [b][n_image][row_vec] or nullptr
    int n_image;
    uint4* out;                       // [b][n_image + len][row_vec]
    void* out_mask;                   // [b][n_image + len] int64 or float32
    int mask_f32;
    int* err;
};

__global__ void __launch_bounds__(256) embed_prompt_kernel(const EmbedParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int out_len = p.n_image + p.len;
    const long long rows = static_cast<long long>(p.b) * out_len;
    for (long long r = warp_global; r < rows; r += n_warps) {
        const int q = static_cast<int>(r / out_len);
        const int pos = static_cast<int>(r - static_cast<long long>(q) * out_len);
        const uint4* src = nullptr;
        long long m = 1;
        if (pos < p.n_image) {
            src = p.image_tokens + (static_cast<size_t>(q) * p.n_image + pos) * p.row_vec;
        } else {
            const size_t i = static_cast<size_t>(q) * p.in_stride + (pos - p.n_image);
            const long long tok = p.input_ids[i];
            m = p.attention_mask[i];
            if (tok >= 0 && tok < p.vocab) src = p.table + static_cast<size_t>(tok) * p.row_vec;
            else if (lane == 0 && p.err) atomicCAS(p.err, 0, kErrTokenRange);
        }
        uint4* dst = p.out + static_cast<size_t>(r) * p.row_vec;
        for (int v = lane; v < p.row_vec; v += 32) dst[v] = src ? __ldg(src + v) : make_uint4(0u, 0u, 0u, 0u);
        if (lane == 0) {
            if (p.mask_f32) static_cast<float*>(p.out_mask)[r] = static_cast<float>(m);
            else static_cast<long long*>(p.out_mask)[r] = m;
        }
    }
}

}  // namespace mpr
