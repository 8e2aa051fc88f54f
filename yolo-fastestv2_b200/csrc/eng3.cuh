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


}  // namespace eng3
}  // namespace yfv2
