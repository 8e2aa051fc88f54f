// Engine v3 building blocks shared by k_blk.cu (stride-1 block chains) and k_tail.cu (small-map chains): the per-warpgroup
// hand-off state and the TMEM / proxy helpers.  See k_blk.cu for the protocol.
#pragma once
#include "common.cuh"
#include "tc.cuh"

namespace yfv2 {
namespace eng3 {

using namespace tc;

struct BPipe {                          // per warpgroup, in shared memory
    uint64_t empty[2];
    uint64_t dfull;
    uint32_t arrivals[2];               // one counter per A buffer: a fast warp may be one chunk ahead of a slow one
};
struct BGrp {
    uint32_t tcol, tlane;               // TMEM address of the group's column block (lane 0 / this warp's lane quarter)
    BPipe* pipe;
    uint32_t chunk;                     // chunks handed over so far
    uint32_t dparity;
    int gtid;                           // 0..127
};

__device__ __forceinline__ uint32_t atom_inc_acq_rel(uint32_t* addr) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(addr)) : "memory");
    return old;
}
// order generic-proxy accesses to shared memory before async-proxy (TMA / tensor core) accesses
__device__ __forceinline__ void publish_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16v(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}


// ---- one 128-row tile per warpgroup: A ring of NB buffers x (KC hi + KC lo) columns at the start of the group's block -----
template <int NB>
__device__ __forceinline__ void st_acquire(BGrp& g) {
    const uint32_t buf = g.chunk % NB, use = g.chunk / NB;
    if (use > 0) mbar_wait(&g.pipe->empty[buf], (use - 1) & 1u);   // the MMAs that last read this buffer have completed
    fence_after_sync();
}
template <int KC, int NB>
__device__ __forceinline__ void st_store(const BGrp& g, const float* a) {
    const uint32_t col = g.tlane + (g.chunk % NB) * (2 * KC);
#pragma unroll
    for (int j = 0; j < KC; j += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            hi[i] = __float_as_uint(a[j + i]) & 0xFFFFE000u;
            lo[i] = __float_as_uint(a[j + i] - __uint_as_float(hi[i]));
        }
        tmem_st8(col + j, hi);
        tmem_st8(col + KC + j, lo);
    }
}
// chunk c of a KP -> NP contraction is in TMEM: the last of the group's four warps to arrive issues its MMAs into the
// accumulator at column dcol of the group's block; `last` also commits the tile's "D full" barrier
template <int KP, int NP, int KC, int NB>
__device__ __forceinline__ void st_hand_off(BGrp& g, int c, uint32_t b_hi, uint32_t b_lo, uint32_t dcol, bool last) {
    wait_st();
    fence_before_sync();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const uint32_t buf = g.chunk % NB;
        const uint32_t old = atom_inc_acq_rel(&g.pipe->arrivals[buf]);
        if ((old & 3u) == 3u) {
            fence_after_sync();
            constexpr uint32_t idesc = make_idesc_tf32(128, NP);
            constexpr uint32_t LBO = 128, SBO = (KP / 4) * 128;
            const uint32_t a_hi = g.tcol + buf * (2 * KC), a_lo = a_hi + KC, d = g.tcol + dcol;
#pragma unroll
            for (int s = 0; s < KC / 8; ++s) {
                const int ks = c * (KC / 8) + s;
                const uint64_t bh = make_b_desc(b_hi + ks * 256, LBO, SBO);
                const uint64_t bl = make_b_desc(b_lo + ks * 256, LBO, SBO);
                mma_tf32_ts(d, a_lo + 8 * s, bh, idesc, ks > 0 ? 1u : 0u);      // small terms first
                mma_tf32_ts(d, a_hi + 8 * s, bl, idesc, 1u);
                mma_tf32_ts(d, a_hi + 8 * s, bh, idesc, 1u);
            }
            mma_commit(&g.pipe->empty[buf]);
            if (last) mma_commit(&g.pipe->dfull);
        }
    }
    __syncwarp();
    ++g.chunk;
}
__device__ __forceinline__ void st_wait_d(BGrp& g) {
    mbar_wait(&g.pipe->dfull, g.dparity);
    g.dparity ^= 1u;
    fence_after_sync();
}

}  // namespace eng3
}  // namespace yfv2
