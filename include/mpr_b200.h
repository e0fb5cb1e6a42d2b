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

#define MPR_ABI_VERSION 2

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

/*
 * Bytes of device scratch a search of this shape needs.  The first bytes of a workspace are control words (grid
 * barrier, tile scheduler, shared admission thresholds) that must be zero when a launch starts: the library zeroes them
 * (one cudaMemsetAsync) the first time it sees a workspace pointer, and every launch leaves them zero again, so a
 * workspace that is reused step after step costs nothing.  A caller that lets the memory be used for anything else —
 * or frees it and may get the same address back from its allocator — must say so with mpr_workspace_invalidate
 * (workspace == NULL: forget every workspace) before the next search.  One workspace serves one stream at a time.
 */
size_t mpr_search_workspace_bytes(mpr_handle_t h, int b, int64_t n_local, int d, int kk);
int mpr_workspace_invalidate(mpr_handle_t h, const void* workspace);

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
 * (tensor or shared memory) — no separate cast launch and no bf16 copy of the queries in HBM.
 * Replaces: cat([image_encoding, text_encoding], 1).float() + cdist + argsort, dataset/VQAFeatureDataset.py:189-197.
 * Available for 64 <= d0+d1 <= 2048 (mpr_search_fused_supported): D <= 512 keeps the q-tile in tensor memory (each
 * epilogue thread converts its own row), larger D — the reference's 1024 — fills the shared-memory q-tile a warp per
 * row.  out_q_bias [b] (may be NULL) receives -0.5*||bf16(q)||^2.
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
 * Candidate exchange over NVLink peer memory (one NVSwitch box).  Every rank allocates one buffer of
 * mpr_exchange_bytes(world, cap) bytes that is mapped into all peers (symmetric memory), zero-filled before first use;
 * cap >= b*kk of any search that will use it.  The exchange itself runs inside mpr_retrieve (see there): per query, the
 * rank's merged local top-kk is stored into every peer's buffer as self-validating 8-byte words {half a key | epoch tag}
 * (plain P2P stores, no fence and no flag), the peers' words of the same query are polled in the rank's own buffer and
 * the `world` lists merged — in the step kernel itself, or (defer_finish) in a small kernel on a side stream.  New in the
 * build: the reference is single-device (main.py:58-61).
 */
size_t mpr_exchange_bytes(int world, int cap);

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
 * One retrieval step, device side — the call RetrievalBank.retrieve_closest_qa_pairs / retrieve_prompt_ids make:
 *   query preparation (cat + optional normalise + bf16 round; dataset/VQAFeatureDataset.py:189-191)
 *   -> bank scan with fused streaming top-(k+skip)  (torch.cdist + torch.argsort[:, s:s+k]; :192-197)
 *   -> merge of the per-CTA partial lists
 *   -> [world > 1] peer-memory exchange of the [b][kk] candidates and merge in rank order
 *   -> [answer_id != NULL] answers of the retrieved rows, majority vote, quantifier bucket, prompt token ids
 *      (:199,215-230 and the tokenizer call of architectures/T5VisionModel.py:153-167)
 * When the scan grid is a single wave (b <= 512 or so) ALL of this is ONE cooperative kernel launch: after a grid
 * barrier the scan kernel's warps finish whole queries.  Larger batches use two launches (scan, tail).
 *
 * Queries: either raw halves q0 [b][d0] (|| q1 [b][d1]) of q_dtype (MPR_SRC_*) — prepared inside the scan kernel — or
 * q_bf16 [b][d] prepared earlier by mpr_bank_build (q0 == NULL).  q_scratch [b][d] bf16 is only needed for raw queries
 * that the scan cannot prepare itself: d > 1024 with several q-tiles (the variant that shares bank tiles between CTA
 * pairs takes prepared queries) and normalised queries with 512 < d <= 1024 beyond 16 of them (hybrid q-tile); without
 * it those shapes take a slower plan.
 * All pointers are device pointers owned by the caller; outputs may be NULL where noted.  The workspace
 * (mpr_search_workspace_bytes) belongs to ONE stream at a time; its control words are (re)zeroed by the library.
 * Sharded search (world > 1) is a COLLECTIVE: every rank must issue the same sequence of calls.
 */
typedef struct mpr_retrieve_args {
    /* queries */
    const void* q0;            /* raw image half [b][d0], or NULL */
    const void* q1;            /* raw text half [b][d1], or NULL  */
    int d0, d1, q_dtype, normalise;
    const uint16_t* q_bf16;    /* prepared queries [b][d] (used when q0 == NULL) */
    uint16_t* q_scratch;       /* [b][d] bf16 or NULL, see above */
    int b;
    /* this rank's bank shard */
    const uint16_t* bank;      /* [n_local][d] bf16 row-major */
    const float* bias;         /* [n_local] = -0.5*|row|^2    */
    int64_t n_local, idx_base;
    int d, kk;                 /* kk = k + skip               */
    /* search results (each may be NULL) */
    uint64_t* out_keys;        /* [b][kk] sortable candidates  */
    float* out_score;          /* [b][kk]                      */
    int32_t* out_idx;          /* [b][kk] global rows, -1 none */
    float* out_q_bias;         /* [b] -0.5*|bf16(q)|^2 (raw queries only) */
    void* workspace;
    size_t workspace_bytes;
    /* multi-GPU exchange: world <= 1 disables it */
    int rank, world, xchg_cap;
    void* const* peer_bufs;    /* HOST array of `world` device pointers (entry `rank` = the local buffer) */
    /* prompt stage: answer_id == NULL disables it */
    int skip;
    const int32_t* answer_id;
    const uint8_t* bucket_lut;
    const int32_t* prefix_ids;
    const int32_t* prefix_off;
    const int32_t* seg_ids;
    const int32_t* seg_off;
    int use_quantifier, pad_id, eos_id, max_len, out_stride;
    int64_t* input_ids;        /* [b][out_stride] */
    int64_t* attention_mask;   /* [b][out_stride] */
    int32_t* out_len;          /* [b] */
    int32_t* maj_answer;       /* [b] */
    int32_t* maj_count;        /* [b] */
    int32_t* bucket;           /* [b] */
    int32_t* ret_answer;       /* [b][k] or NULL */
    int32_t* status;           /* device word, 0 = ok, MPR_STATUS_* otherwise (may be NULL) */
    /* world > 1 only.  Non-zero: the step kernel only PUSHES this rank's candidates to the peers; collecting the peers'
     * candidates, the merge, the vote and every output of the step are left to a small second kernel that the library
     * queues on a stream of its own, so that the NVLink latency of the exchange (~10 us) hides under whatever follows
     * on `stream` — normally the next step's scan.  The outputs are then NOT ordered before later work on `stream`:
     * call mpr_retrieve_join(h, some_stream) before consuming them there (mpr_retrieve_host does it for its own
     * device-to-host copy).  At most two deferred steps are left unfinished at any time (the library makes step j+2
     * wait for the finish of step j).  Ignored for batches that take the two-launch path. */
    int defer_finish;
} mpr_retrieve_args;

#define MPR_STATUS_XCHG_TIMEOUT 201 /* a peer rank did not deliver its candidates in time; local results returned */

int mpr_retrieve(mpr_handle_t h, const mpr_retrieve_args* a, void* stream);

/* Makes `stream` wait for every deferred finish (mpr_retrieve_args.defer_finish) queued so far on this handle. */
int mpr_retrieve_join(mpr_handle_t h, void* stream);

/*
 * The same step for HOST-resident inputs and outputs (what a data-loader thread hands over and what the T5 side of
 * prepare_input consumes): copies the pinned host query halves and the host-tokenised prefix CSR to the device
 * staging buffers named in `a` / `io`, runs mpr_retrieve, and copies ONE contiguous result block back:
 *   d_out/h_out layout = whatever the caller laid out behind a->input_ids ... (the block [d_out, d_out + out_bytes)
 *   must contain every output pointer of `a` that the caller wants on the host).
 * Everything is asynchronous on `stream`; with sync != 0 the call returns after the results are in h_out.
 * h_q0 == NULL: the queries are already on the device (a->q0 / a->q_bf16 as in mpr_retrieve).
 */
typedef struct mpr_host_io {
    const void* h_q0;          /* pinned host [b][d0] of a->q_dtype, copied to a->q0 */
    const void* h_q1;          /* pinned host [b][d1] or NULL, copied to a->q1       */
    const int32_t* h_prefix_ids;
    int n_prefix_ids;          /* copied to a->prefix_ids                            */
    const int32_t* h_prefix_off; /* [b+1], copied to a->prefix_off                   */
    const void* d_out;         /* device result block                                */
    void* h_out;               /* pinned host result block                           */
    size_t out_bytes;
    int sync;
    /* Optional copy streams (cudaStream_t; NULL = everything on `stream`).  With stream_in the host-to-device copies
     * run on it and the step on `stream` waits for them; with stream_out the device-to-host copy runs on it after the
     * step.  A caller that keeps two steps in flight (staging and result buffers alternating) thereby overlaps step
     * i+1's input copy and step i-1's result copy with step i's kernel.  sync then waits on stream_out.  The caller
     * owns the ordering of buffer reuse: a staging / result buffer may be handed to a new call only after the call
     * that last used it has completed (e.g. an event recorded on stream_out). */
    void* stream_in;
    void* stream_out;
} mpr_host_io;

int mpr_retrieve_host(mpr_handle_t h, const mpr_retrieve_args* a, const mpr_host_io* io, void* stream);

/* Seconds a rank waits for a peer's candidates inside a sharded search before giving up (default 60). */
int mpr_set_exchange_timeout(mpr_handle_t h, double seconds);

/*
 * Kernel 5 — prompt token ids -> T5 input embeddings with the image tokens prepended, forward only.
 * Replaces: T5_model.shared(input_ids), torch.ones image mask, and the two torch.cat calls of
 *           architectures/T5VisionModel.py:169-176.
 * input_ids / attention_mask: int64 [b][in_stride] on the device (kernel 3's outputs), first `len` columns used.
 * table [vocab][hidden] and image_tokens [b][n_image][hidden] (may be NULL with n_image = 0: "only use question",
 * :178-180) share table_dtype (MPR_SRC_*).  out_embeds [b][n_image+len][hidden] in the same dtype;
 * out_mask [b][n_image+len] is float32 (mask_f32 != 0: what the reference's cat of a float ones tensor with the int64
 * tokenizer mask promotes to) or int64.  An id outside [0, vocab) zero-fills its row and sets the device error word.
 */
int mpr_embed_prompt(mpr_handle_t h, const int64_t* input_ids, const int64_t* attention_mask, int b, int len,
                     int in_stride, const void* table, int table_dtype, int vocab, int hidden,
                     const void* image_tokens, int n_image, void* out_embeds, void* out_mask, int mask_f32,
                     void* stream);

/*
 * Tuning aid: with MPR_DEBUG_COUNTERS=1 in the environment at mpr_create the scan kernel counts, since the last call,
 * [0] candidates admitted past the thresholds, [1] warp-level list flushes, [2] 8-score groups that took the admission
 * path, [3] list replacements, [4] warp-tiles processed, [5] threshold refreshes that found a shared bound.  Synchronises
 * the device; returns zeros when the counters are off.
 */
int mpr_debug_counters(mpr_handle_t h, uint64_t* out8);
/* With the same switch: out[24*cta + k] = globaltimer (ns) at event k of CTA cta in the LAST scan launch (k: 0 entry,
 * 1 q-tile ready, 2 producer done, 3/5 epilogue group 0/1 left its tile loop, 4/6 its partial lists written, 7 CTA done,
 * 8 past the grid barrier, 9 tail done, 10 second tile's data arrived, 11/12 first and 13/14 fourth tile consumed,
 * 15 first tile's accumulator ready, 16-19 its four 32-row chunks done, 20 its buffer released, 21 second tile's bias staged). */
int mpr_debug_timeline(mpr_handle_t h, uint64_t* out, int n_ctas);
/* With the same switch: for each of the last 64 scan launches (entry = launch number mod 64) out128[2e] = 2^63 - globaltimer
 * of its first CTA's first instruction, out128[2e+1] = globaltimer of its last CTA's last instruction; *next_seq = number
 * of launches so far.  Clears the ring.  Shows the gap between back-to-back launches. */
int mpr_debug_launch_ring(mpr_handle_t h, uint64_t* out128, unsigned* next_seq);

/* Kernel launches issued by the last mpr_retrieve / mpr_retrieve_host on this handle (bench bookkeeping). */
int mpr_last_launch_count(mpr_handle_t h);

/*
 * Host-side token cache (no CUDA): builds the per-query prefix CSR (prefix_ids / prefix_off of mpr_retrieve_args) from
 * cached tokenisations of space-separated chunks.  Replaces the per-batch tokenizer call of
 * architectures/T5VisionModel.py:161-167 for everything that was seen before; the caller keeps the tokenizer and
 * registers unseen chunks with mpr_token_cache_put.
 *   put: registers m chunks at once — chunk j = chunks[chunk_off[j] .. chunk_off[j+1]), its tokens ids[ids_off[j] ..
 *   ids_off[j+1]).
 *   assemble: the n texts lie back to back in `texts` (text i = bytes [text_off[i], text_off[i+1]), ASCII).  Row i of the
 *   CSR = the head head_index[i] of the head table (head_ids / head_off: tokens of "Answer the {task} question:" per
 *   distinct task; head_index may be NULL) followed by the tokens of every maximal run of non-space bytes of text i.
 *   out_off [n+1]; out_ids holds out_cap ids (MPR_EWORKSPACE if too small); *longest = the longest row.  Chunks not in the
 *   cache are reported as (byte offset in `texts`, length) pairs in missing [max_missing][2] and counted in *n_missing;
 *   the output is complete only when *n_missing == 0.
 * All functions are thread-safe (one mutex per cache) and never touch the device.
 */
typedef struct mpr_token_cache* mpr_token_cache_t;
int mpr_token_cache_create(mpr_token_cache_t* out);
int mpr_token_cache_destroy(mpr_token_cache_t c);
int64_t mpr_token_cache_size(mpr_token_cache_t c);
int mpr_token_cache_clear(mpr_token_cache_t c);
int mpr_token_cache_put(mpr_token_cache_t c, int m, const char* chunks, const int32_t* chunk_off, const int32_t* ids,
                        const int32_t* ids_off);
int mpr_token_cache_assemble(mpr_token_cache_t c, int n, const char* texts, const int32_t* text_off,
                             const int32_t* head_ids, const int32_t* head_off, const int32_t* head_index,
                             int32_t* out_ids, int64_t out_cap, int32_t* out_off, int32_t* missing, int max_missing,
                             int32_t* n_missing, int32_t* longest);

/*
 * Debug / test aid: the full [b][n_local] score matrix through the SAME tcgen05 pipeline as mpr_search_topk
 * (the kernel is instantiated with a dump epilogue).  Small shapes only.
 */
int mpr_debug_scores(mpr_handle_t h, const uint16_t* q, int b, const uint16_t* bank, const float* bias,
                     int64_t n_local, int d, float* scores, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Measurement aid for bench.py's roofline: between begin and end every scan-kernel launch (kernel 2 including its
 * fused tail, not a separate tail launch) is bracketed by a cudaEvent pair on the launching stream.  mpr_profile_end synchronises on the last event and
 * returns the summed device time and the number of launches.  At most max_launches launches are recorded.
 */
int mpr_profile_begin(mpr_handle_t h, int max_launches);
int mpr_profile_end(mpr_handle_t h, float* total_ms, int* n_launches);
/* Device time of the i-th launch recorded by the last begin/end pair (valid after mpr_profile_end). */
int mpr_profile_launch_ms(mpr_handle_t h, int i, float* ms);

/* Launch geometry the library would use for a shape (for bench/roofline bookkeeping). */
int mpr_search_plan(mpr_handle_t h, int b, int64_t n_local, int d, int kk, int* n_ctas, int* n_splits,
                    int* n_qtiles, int* n_stages, int* smem_bytes);

/*
 * The same planning WITHOUT a device (host-only; what the library would do on a GPU with `num_sms` SMs and default
 * tuning): out[0] = CTAs, [1] = CTAs per q-tile, [2] = q-tiles, [3] = ring stages, [4] = dynamic shared memory bytes,
 * [5] = queries per q-tile, [6] = 64-wide K sub-chunks per ring stage, [7] = epilogue groups, [8] = q-tile in tensor
 * memory, [9] = hybrid q-tile, [10] = register lists, [11] = pending-candidate slots, [12] = rows of a q-tile slab in
 * shared memory, [13] = workspace bytes (low 31 bits), [14] = threshold slots, [15] = replica words per slot.
 * Lets the planner's invariants be tested on a machine without a GPU.
 */
int mpr_plan_host(int num_sms, int b, int64_t n_local, int d, int kk, int* out16);

#ifdef __cplusplus
}
#endif
#endif /* MPR_B200_H */
