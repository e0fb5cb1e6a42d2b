// Candidate keys and the warp-cooperative sorted-list insert shared by the scan epilogue and the merge kernel.
//
// A candidate (score, bank row) is one u64:  hi = order-preserving image of the fp32 score,
// lo = ~row.  Unsigned comparison of keys is then "higher score first, lower row on ties" — the
// deterministic tie rule the build adopts for the reference's unstable argsort
// (/root/reference/dataset/VQAFeatureDataset.py:195,197).  key == 0 means "empty slot".
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mpr {

constexpr unsigned kFullMask = 0xFFFFFFFFu;

__host__ __device__ __forceinline__ uint32_t f32_to_ordered(uint32_t bits) {
    return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t ordered_to_f32(uint32_t o) {
    return (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
    return (static_cast<uint64_t>(f32_to_ordered(__float_as_uint(score))) << 32) | static_cast<uint64_t>(~row);
}
__device__ __forceinline__ float key_score(uint64_t key) {
    return __uint_as_float(ordered_to_f32(static_cast<uint32_t>(key >> 32)));
}
__device__ __forceinline__ int32_t key_row(uint64_t key) {
    return key == 0 ? -1 : static_cast<int32_t>(~static_cast<uint32_t>(key));
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(kFullMask, static_cast<uint32_t>(v), src);
    uint32_t hi = __shfl_sync(kFullMask, static_cast<uint32_t>(v >> 32), src);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int delta) {
    uint32_t lo = __shfl_up_sync(kFullMask, static_cast<uint32_t>(v), delta);
    uint32_t hi = __shfl_up_sync(kFullMask, static_cast<uint32_t>(v >> 32), delta);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

// One sorted (descending) list of up to 32 keys spread over the lanes of a warp: lane i holds element i.
// Inserts `key` (same value in every lane; caller guarantees key > element[kk-1]) and returns the
// updated element for this lane.  Elements at lanes >= kk are scratch.
__device__ __forceinline__ uint64_t warp_list_insert(uint64_t elem, uint64_t key, int lane) {
    const unsigned beats = __ballot_sync(kFullMask, key > elem);   // a suffix of the lanes (list is sorted)
    const int pos = __ffs(beats) - 1;                              // first element the key outranks
    const uint64_t up = shfl_up_u64(elem, 1);
    return lane < pos ? elem : (lane == pos ? key : up);
}

}  // namespace mpr
