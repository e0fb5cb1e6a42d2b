// Candidate exchange over NVLink peer memory — the multi-GPU half of the top-k, without NCCL.
//
// Every rank owns one "exchange buffer" of identical layout, mapped into every peer (symmetric memory):
//
//   byte 0      u32 epoch          last exchange this rank has fully consumed
//   byte 4      u32 done           warps of the current merge that have finished
//   byte 64     u32 flag[64]       flag[r] = newest epoch whose candidates rank r has delivered INTO THIS buffer
//   byte 1024   u64 slot[2][world][cap]   candidate lists, double-buffered on epoch parity
//
// push  (1 CTA):  e = epoch + 1; store my [b][kk] keys into slot[e&1][my_rank] of EVERY rank's buffer (plain P2P stores),
//                 __threadfence_system, then st.release.sys flag[my_rank] = e on every rank.
// merge (1 warp per query): wait (ld.acquire.sys) until flag[r] >= e for all r, k-way merge of slot[e&1][0..world) in
//                 rank order (= global row order, so ties still resolve to the lower row), last warp publishes epoch = e.
//
// Why two slots are enough: a rank can start exchange e+1 (writing slot[(e+1)&1]) while a slow peer still reads
// slot[e&1], but it cannot reach e+2 before that peer has delivered its own e+1 flags, which it does only after its
// merge of e.  The epoch lives in device memory, so the whole chain is CUDA-graph capturable.
// The waiting kernel only ever waits on OTHER GPUs (or on an earlier kernel of its own stream), never on a kernel that
// must be co-scheduled on the same device.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <math_constants.h>

#include "ptx.cuh"
#include "topk_key.cuh"

namespace mpr {

constexpr int kXchgMaxWorld = 16;
constexpr int kXchgFlagOff = 64;
constexpr int kXchgSlotOff = 1024;
constexpr int kErrXchgTimeout = 201;

struct XchgPeers {
    unsigned char* buf[kXchgMaxWorld];
};

__host__ __device__ inline size_t xchg_bytes(int world, int cap) {
    return static_cast<size_t>(kXchgSlotOff) + 2ull * world * cap * sizeof(uint64_t);
}

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_cg_u64(const uint64_t* p) {   // L2-coherent load: peers write this memory
    uint64_t v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(1024)
xchg_push_kernel(const uint64_t* __restrict__ local_keys, int n_keys, int rank, int world, int cap, XchgPeers peers) {
    const uint32_t e = *reinterpret_cast<volatile uint32_t*>(peers.buf[rank]) + 1u;
    const size_t slot = (static_cast<size_t>(e & 1u) * world + rank) * cap;
    for (int p = 0; p < world; ++p) {
        uint64_t* dst = reinterpret_cast<uint64_t*>(peers.buf[p] + kXchgSlotOff) + slot;
        for (int i = threadIdx.x; i < n_keys; i += blockDim.x) dst[i] = local_keys[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world)
        st_release_sys_u32(reinterpret_cast<uint32_t*>(peers.buf[threadIdx.x] + kXchgFlagOff) + rank, e);
}

__global__ void __launch_bounds__(128)
xchg_merge_kernel(unsigned char* __restrict__ my_buf, int world, int cap, int b, int kk, uint64_t* __restrict__ out_keys,
                  float* __restrict__ out_score, int32_t* __restrict__ out_idx, int* err) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= b) return;
    uint32_t* ctrl = reinterpret_cast<uint32_t*>(my_buf);
    const uint32_t e = *reinterpret_cast<volatile uint32_t*>(ctrl) + 1u;

    // ---- wait for every rank's delivery of exchange e
    if (lane < world) {
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(my_buf + kXchgFlagOff) + lane;
        if (static_cast<int32_t>(ld_acquire_sys_u32(flag) - e) < 0) {
            const uint64_t t0 = ptx::globaltimer_ns();
            uint32_t polls = 0;
            while (static_cast<int32_t>(ld_acquire_sys_u32(flag) - e) < 0) {
                if ((++polls & 0xFFu) == 0 && ptx::globaltimer_ns() - t0 > 4000000000ull) {
                    if (err) atomicCAS(err, 0, kErrXchgTimeout);
                    __threadfence_system();
                    __trap();
                }
            }
        }
    }
    __syncwarp();

    // ---- k-way merge over the world lists of this query (rank-major walk, early exit as in merge_topk_kernel)
    const uint64_t* slots = reinterpret_cast<const uint64_t*>(my_buf + kXchgSlotOff) +
                            static_cast<size_t>(e & 1u) * world * cap;
    uint64_t elem = 0ull, kth = 0ull;
    for (int i = 0; i < kk; ++i) {
        bool admitted = false;
        for (int l0 = 0; l0 < world; l0 += 32) {
            const int l = l0 + lane;
            const uint64_t key = l < world ? ld_cg_u64(slots + static_cast<size_t>(l) * cap + static_cast<size_t>(q) * kk + i)
                                           : 0ull;
            unsigned pending = __ballot_sync(kFullMask, key > kth);
            while (pending) {
                const int src = __ffs(pending) - 1;
                pending &= pending - 1;
                const uint64_t cand = shfl_u64(key, src);
                if (cand > kth) {
                    elem = warp_list_insert(elem, cand, lane);
                    kth = shfl_u64(elem, kk - 1);
                    admitted = true;
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

    // ---- the last warp to finish publishes the epoch (stream order makes it visible to the next push)
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        const uint32_t done = atomicAdd(ctrl + 1, 1u);
        if (done == static_cast<uint32_t>(b) - 1u) {
            ctrl[1] = 0u;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(ctrl) = e;
        }
    }
}

}  // namespace mpr
