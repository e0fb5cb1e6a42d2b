/*
 * mpr_b200.h — C ABI of the B200-native retrieval hot path for MPR_Gen (tossowski/MultimodalPromptRetrieval).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.  The reference has no FFI of its
 * own (it is pure Python calling torch ops), so every entry point cites the reference lines it replaces; the
 * Python host class `multimodalpromptretrieval_b200.RetrievalBank` binds these with ctypes and mirrors
 * VQADataset.create_retrieval_dataset / retrieve_closest_qa_pairs.
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on failure (MPR_E*); mpr_last_error(h) gives the text.
 *   - all device pointers are owned by the caller (torch); the library allocates nothing persistent on the device
 *     except a 4-byte error word inside the handle.
 *   - all launches are asynchronous on the cudaStream_t passed in (as void*); no device synchronisation inside;
 *     every call is CUDA-graph capturable.
 *   - there is NO CPU fallback: without an sm_100 device mpr_create fails.
 */
#ifndef MPR_B200_H
#define MPR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPR_ABI_VERSION 1

enum {
    MPR_OK = 0,
    MPR_EINVAL = -1,    /* bad argument (alignment, D % 64, k range, null pointer ...) */
    MPR_ECUDA = -2,     /* a CUDA runtime / driver call failed                        */
    MPR_EARCH = -3,     /* device is not sm_100                                       */
    MPR_EWORKSPACE = -4 /* workspace too small                                        */
};

enum { MPR_SRC_F32 = 0, MPR_SRC_F16 = 1, MPR_SRC_BF16 = 2 };

#define MPR_MAX_KK 32 /* k + skip (skip = 1 in the training phase) */

typedef struct mpr_context* mpr_handle_t;

int mpr_abi_version(void);

/* One handle per (process, device).  Fails with MPR_EARCH unless the device is compute capability 10.x. */
int mpr_create(int device, mpr_handle_t* out);
int mpr_destroy(mpr_handle_t h);
const char* mpr_last_error(mpr_handle_t h);
/* Reads (and clears) the device-side error word: non-zero = the code of a pipeline barrier that starved. */
int mpr_device_error(mpr_handle_t h, int* code);

/*
 * Kernel 1 — row preparation.  out[r] = bf16(concat(src0[r], src1[r]) / (normalise ? ||.||2 : 1)),
 * bias[r] = -0.5 * ||out[r]||^2 (fp32, computed from the rounded values).
 * Replaces: torch.cat([image_encoding, text_encoding], 1) ... torch.cat(all_embeddings, 0).float()
 *           dataset/VQAFeatureDataset.py:146-148,159,179 (bank) and :189-191 (queries),
 *           and the per-call ||b||^2 recomputation inside torch.cdist (:192).
 * src1 may be NULL (d1 = 0).  d0, d1 multiples of 8; (d0 + d1) % 64 == 0 and <= 2048; pointers 16-byte aligned.
 */
int mpr_bank_build(mpr_handle_t h, const void* src0, int d0, const void* src1, int d1, int src_dtype, int64_t n,
                   int normalise, uint16_t* out_bf16, float* out_bias, void* stream);

/* Bytes of device scratch mpr_search_topk needs for this shape. */
size_t mpr_search_workspace_bytes(mpr_handle_t h, int b, int64_t n_local, int d, int kk);

/*
 * Kernel 2 (+ kernel 4 over the bank splits) — score every query against this shard and keep the best kk rows.
 * score = <q, bank[r]> + bias[r]; ordering: score descending, global row ascending on ties.
 * Replaces: torch.cdist(combined, retrieval_embeddings) + torch.argsort(dist, 1)[:, 0:kk]
 *           dataset/VQAFeatureDataset.py:192-197  (and torch.sort(...).values[:, :k] at :243).
 * q [b][d] bf16, bank [n_local][d] bf16 row-major, bias [n_local] fp32.  idx_base = global index of shard row 0.
 * Outputs (each may be NULL): out_keys [b][kk] u64 sortable candidates (what the cross-GPU allgather carries),
 * out_score [b][kk] fp32, out_idx [b][kk] int32 global rows (-1 = fewer than kk rows exist).
 */
int mpr_search_topk(mpr_handle_t h, const uint16_t* q, int b, const uint16_t* bank, const float* bias,
                    int64_t n_local, int64_t idx_base, int d, int kk, uint64_t* out_keys, float* out_score,
                    int32_t* out_idx, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Kernel 2 with the query preparation fused in: the raw CLIP outputs src0 [b][d0] (‖ src1 [b][d1])
 * are concatenated, optionally L2-normalised and rounded to bf16 by the scan kernel itself while it loads the q-tile
 * into tensor memory — no separate cast launch and no bf16 copy of the queries in HBM.
 * Replaces: cat([image_encoding, text_encoding], 1).float() + cdist + argsort, dataset/VQAFeatureDataset.py:189-197.
 * Available when the q-tile lives in tensor memory: 64 <= d0+d1 <= 512 (mpr_search_fused_supported); the reference's
 * D = 1024 uses mpr_bank_build + mpr_search_topk.  out_q_bias [b] (may be NULL) receives -0.5*||bf16(q)||^2.
 * With normalise = 0 the results are bit-identical to mpr_bank_build + mpr_search_topk.
 */
int mpr_search_fused_supported(mpr_handle_t h, int d);
int mpr_search_topk_fused(mpr_handle_t h, const void* src0, int d0, const void* src1, int d1, int src_dtype,
                          int normalise, int b, const uint16_t* bank, const float* bias, int64_t n_local,
                          int64_t idx_base, int kk, uint64_t* out_keys, float* out_score, int32_t* out_idx,
                          float* out_q_bias, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Kernel 4 — merge n_lists sorted candidate lists per query: in_keys [n_lists][b][kk] -> top-kk.
 * Used after the NCCL allgather of every rank's out_keys (n_lists = world size).  New in the build: the reference
 * is single-device (main.py:58-61).
 */
int mpr_merge_topk(mpr_handle_t h, const uint64_t* in_keys, int n_lists, int b, int kk, uint64_t* out_keys,
                   float* out_score, int32_t* out_idx, void* stream);

/*
 * Candidate exchange over NVLink peer memory (alternative to NCCL all-gather + mpr_merge_topk on one NVSwitch box;
 * csrc/exchange.cuh).  Every rank allocates one buffer of mpr_exchange_bytes(world, cap) bytes that is mapped into all
 * peers (symmetric memory), zero-filled before first use.  peer_bufs is a HOST array of `world` device pointers, entry r
 * = rank r's buffer as addressable from this device (entry `rank` = the local buffer).  cap >= b*kk.
 *   mpr_exchange_push : store this rank's [b][kk] keys into every rank's buffer and raise the delivery flags.
 *   mpr_exchange_merge: wait for all `world` deliveries, merge them (score desc, row asc) into the global top-kk.
 * Both are single launches on `stream`, keep their epoch in device memory and are CUDA-graph capturable.  New in the
 * build: the reference is single-device (main.py:58-61).
 */
size_t mpr_exchange_bytes(int world, int cap);
int mpr_exchange_push(mpr_handle_t h, const uint64_t* local_keys, int b, int kk, int rank, int world,
                      void* const* peer_bufs, int cap, void* stream);
int mpr_exchange_merge(mpr_handle_t h, void* my_buf, int world, int cap, int b, int kk, uint64_t* out_keys,
                       float* out_score, int32_t* out_idx, void* stream);

/*
 * Kernel 3 — retrieved rows -> answers -> majority vote -> quantifier bucket -> prompt token ids.
 * Replaces: dataset/VQAFeatureDataset.py:199,215-230 and the tokenizer call of
 *           architectures/T5VisionModel.py:153-167 (padding="longest", truncation to max_len, </s> appended).
 * idx [b][kk]; the votes are ranks skip..kk-1 (skip = 1 reproduces the training-phase slice [:, 1:1+k]).
 * Segment table (CSR seg_ids/seg_off): 0 "believe the answer is", 1 "most frequent answer is", 2..7 the six
 * buckets, 8+a answer a.  prefix CSR: per-query tokens of "Answer the {task} question: " + question + "I"|"The".
 * bucket_lut [(k+1)*(k+1)] with lut[n_votes*(k+1)+max_count] = int(max_count / n_votes * 5) (host, float64).
 * input_ids / attention_mask: int64 [b][out_stride], padded with pad_id / 0; out_len[b] includes </s>.
 */
int mpr_prompt_gather(mpr_handle_t h, const int32_t* idx, int b, int kk, int skip, const int32_t* answer_id,
                      const uint8_t* bucket_lut, const int32_t* prefix_ids, const int32_t* prefix_off,
                      const int32_t* seg_ids, const int32_t* seg_off, int use_quantifier, int pad_id, int eos_id,
                      int max_len, int out_stride, int64_t* input_ids, int64_t* attention_mask, int32_t* out_len,
                      int32_t* maj_answer, int32_t* maj_count, int32_t* bucket, int32_t* ret_answer, void* stream);

/*
 * Debug / test aid: the full [b][n_local] score matrix through the SAME tcgen05 pipeline as mpr_search_topk
 * (the kernel is instantiated with a dump epilogue).  Small shapes only.
 */
int mpr_debug_scores(mpr_handle_t h, const uint16_t* q, int b, const uint16_t* bank, const float* bias,
                     int64_t n_local, int d, float* scores, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Measurement aid for bench.py's roofline: between begin and end every scan-kernel launch (kernel 2 only, not the
 * merge) is bracketed by a cudaEvent pair on the launching stream.  mpr_profile_end synchronises on the last event and
 * returns the summed device time and the number of launches.  At most max_launches launches are recorded.
 */
int mpr_profile_begin(mpr_handle_t h, int max_launches);
int mpr_profile_end(mpr_handle_t h, float* total_ms, int* n_launches);
/* Device time of the i-th launch recorded by the last begin/end pair (valid after mpr_profile_end). */
int mpr_profile_launch_ms(mpr_handle_t h, int i, float* ms);

/* Launch geometry the library would use for a shape (for bench/roofline bookkeeping). */
int mpr_search_plan(mpr_handle_t h, int b, int64_t n_local, int d, int kk, int* n_ctas, int* n_splits,
                    int* n_qtiles, int* n_stages, int* smem_bytes);

#ifdef __cplusplus
}
#endif
#endif /* MPR_B200_H */
