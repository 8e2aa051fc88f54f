// tcgen05 (5th-gen tensor core) building blocks for the 1x1 "pointwise" contractions, sm_100a only.
//
// Scheme: error-compensated TF32 ("3xTF32").  fp32 operand a is split as a = hi + lo with hi = rna_tf32(a),
// lo = rna_tf32(a - hi); the product A*B is accumulated as Ahi*Bhi + Alo*Bhi + Ahi*Blo in the fp32 TMEM
// accumulator (the dropped Alo*Blo term is ~2^-22 relative).  SURVEY 7 hard part 1: single-pass TF32 misses
// the 1e-4 parity bar by 300x, the 3-pass split passes it.
//
// Data flow per 128-pixel tile, executed by one warpgroup (4 warps = 128 threads, thread = pixel = TMEM lane):
//   registers (this pixel's K inputs) --split--> tcgen05.st --> TMEM A_hi / A_lo   (A operand lives in TMEM)
//   weights: pre-split, pre-tiled in the UMMA K-major no-swizzle canonical layout in shared memory (B operand)
//   one elected thread: 3*K/8 tcgen05.mma.kind::tf32 (M=128, N=NP, K=8 each) --> TMEM D, tcgen05.commit -> mbarrier
//   all threads: mbarrier wait, tcgen05.ld D --> registers --> epilogue (BN scale/shift, ReLU, store)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace yfv2 {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM allocation (one warp, .sync.aligned) -------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- mbarrier ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Bounded spin: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
#pragma unroll 1        // (left to itself nvcc unrolls this spin 64x at every call site: ~2 KB of SASS each)
    for (uint32_t it = 0; it < (1u << 26); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMA bulk copy (UBLKCP): contiguous global -> shared, completion counted in bytes on an mbarrier ------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// named barrier for one warpgroup (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void group_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---- descriptors ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave") canonical layout:
//   core matrix = 8 rows (N) x 16 bytes (4 tf32 of K), 128 contiguous bytes;
//   LBO = byte distance between the two core matrices an MMA (K=8) touches along K;
//   SBO = byte distance between consecutive 8-row groups along N.
// Bits: [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1 (Blackwell), [61,64) layout=0.
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// Instruction descriptor: D fp32 (c_format=1 @4), A/B tf32 (format 2 @7, @10), A and B K-major, N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]     (A from tensor memory, B via descriptor)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// ---- TMEM <-> registers (32x32b: thread t of warp w touches lane 32*(w%4)+t, N consecutive columns) -------------------
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// Split 8 fp32 values and store them as 8 columns of A_hi (at col) and A_lo (at col + lo_off) for this thread's lane.
// hi keeps the top 19 bits (sign, exponent, 10 mantissa bits) so it is exactly a tf32 value; lo = a - hi is exact in fp32
// (|lo| < 2^-10 |a|) and is handed over as is — the tensor core reads its top 19 bits, leaving a residual below
// 2^-20 |a| per operand.  Two instructions per element instead of two emulated cvt.rna.tf32.
__device__ __forceinline__ void store_a8(uint32_t taddr_hi, uint32_t lo_off, const float (&a)[8]) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        hi[i] = __float_as_uint(a[i]) & 0xFFFFE000u;
        lo[i] = __float_as_uint(a[i] - __uint_as_float(hi[i]));
    }
    tmem_st8(taddr_hi, hi);
    tmem_st8(taddr_hi + lo_off, lo);
}

// Packed B operand for one PW layer (floats): Bhi[NP*KP] | Blo[NP*KP] | scale[NP] | shift[NP], tile order [n/8][k/4][n%8][k%4].
__host__ __device__ constexpr int tc_round(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ constexpr int tc_pack_floats(int K, int N) { return 2 * tc_round(N, 16) * tc_round(K, 8) + 2 * tc_round(N, 16); }
__host__ __device__ constexpr int tc_b_index(int n, int k, int KP) { return (n >> 3) * (KP * 8) + (k >> 2) * 32 + (n & 7) * 4 + (k & 3); }

// Issue the 3*KC/8 MMAs of one K-chunk of a PW for one 128-row tile.  One thread calls this.
//   a_hi / a_lo / d: TMEM addresses (lane 0) of the operand column blocks (the chunk's KC columns each);
//   b_hi / b_lo: shared addresses of the FULL packs (KPFULL columns per row); chunk_k0 = first K index of the chunk.
template <int KC, int NP, int KPFULL>
__device__ __forceinline__ void issue_pw(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi_smem, uint32_t b_lo_smem,
                                         int chunk_k0, bool accumulate_first) {
    constexpr uint32_t idesc = make_idesc_tf32(128, NP);
    constexpr uint32_t LBO = 128, SBO = (KPFULL / 4) * 128;
#pragma unroll
    for (int s = 0; s < KC / 8; ++s) {
        const uint32_t koff = (uint32_t)(chunk_k0 / 8 + s) * 256;
        const uint64_t bh = make_b_desc(b_hi_smem + koff, LBO, SBO);
        const uint64_t bl = make_b_desc(b_lo_smem + koff, LBO, SBO);
        mma_tf32_ts(d, a_lo + 8 * s, bh, idesc, (s > 0 || accumulate_first) ? 1u : 0u);   // small terms first
        mma_tf32_ts(d, a_hi + 8 * s, bl, idesc, 1u);
        mma_tf32_ts(d, a_hi + 8 * s, bh, idesc, 1u);
    }
}

}  // namespace tc
}  // namespace yfv2
