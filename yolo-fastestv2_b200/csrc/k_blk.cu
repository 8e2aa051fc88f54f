// ShuffleNetV2 blocks, engine v3 (reference model/backbone/shufflenetv2.py:19-63).
//
// Same arithmetic as k_tcnet.cu's tc_s1 / tc_s2 kernels (3xTF32 pointwise on tcgen05 with the A operand in TMEM,
// depthwise 3x3 on the CUDA cores, BN scale/shift + ReLU in the epilogue: results are bit-identical), rebuilt around
// what the round-1 ncu captures showed (profiles/r2_s1_48_before.txt): 12.6 issued instructions per useful FFMA
// (table lookups, 64-bit address math, predicated loads, per-8-channel hand-offs with a shared atomic each, a
// division-heavy zero-fill per item, three CTA barriers per half-image item, 40 KB of weights pulled through
// registers by every CTA).  Here:
//   * one persistent CTA per SM owns WHOLE IMAGES whenever the pointwise-1 output of an image fits in shared
//     memory (T: K planes with a one-pixel zero frame; 352x352 input: all of stage 2 and stage 3), so there is no
//     band halo to recompute, the zero frame is written once per kernel, and consecutive stride-1 blocks of a stage
//     are CHAINED inside one launch: block b+1 of an image is run by the CTA that ran block b, its input comes out
//     of L2, and the weights of block b+1 arrive by TMA bulk copy while block b computes;
//   * the A operand goes to the tensor core in chunks of KC channels (16 or the whole K) instead of 8;
//   * every per-channel plane offset is a kernel-parameter constant (no table lookups / multiplies in the loops),
//     weights reach shared memory by cp.async.bulk (UBLKCP), invalid lanes are clamped instead of predicated.
#include <cstdlib>

#include "eng3.cuh"

namespace yfv2 {
namespace {

using namespace tc;
using namespace eng3;

constexpr int kMaxChain = 7;            // stride-1 blocks per launch (stage 3 has seven)
constexpr int kMaxK = 48;               // branch width handled here (K = 96 blocks: k_tcnet.cu)

// ---------------------------------------------------------------------------------------------------------
// Pixel pairs.  A thread owns TWO vertically adjacent pixels (rows 2j and 2j+1 of the band, same column): the same TMEM lane
// of two M=128 tiles, so a warpgroup drives two accumulators.  The depthwise stencil then needs 12 shared-memory loads for
// two outputs instead of 18 and loads the channel's weights once (the kernel is bound by shared-memory wavefronts: ncu on
// the one-pixel version showed 4.2 MIO-throttle stalls per issue and 38 M wavefronts per launch, 13 M of them bank
// conflicts from the padded rows).  TMEM columns of a group: NB ring buffers of [tile0 hi KC | tile0 lo KC | tile1 hi KC |
// tile1 lo KC], then D0[NP], D1[NP].
// ---------------------------------------------------------------------------------------------------------
template <int KC, int NB>
__device__ __forceinline__ void acquire_buf(BGrp& g) {
    const uint32_t buf = g.chunk % NB, use = g.chunk / NB;
    if (use > 0) mbar_wait(&g.pipe->empty[buf], (use - 1) & 1u);
    fence_after_sync();
}
// KCS (= KC / siblings) channel values of one of this thread's two pixels -> tf32 hi / lo columns [koff, koff + KCS) of the
// current A buffer.
template <int KC, int NB, int KCS = KC>
__device__ __forceinline__ void store_a(const BGrp& g, int tile, const float* a, int koff = 0) {
    const uint32_t col = g.tlane + (g.chunk % NB) * (4 * KC) + tile * (2 * KC) + koff;
#pragma unroll
    for (int j = 0; j < KCS; j += 8) {
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
// Chunk c (of KP / KC) of both tiles is in TMEM: the last of the group's four warps to get here issues the MMAs.
template <int KP, int NP, int KC, int NB, int SIB = 1>
__device__ __forceinline__ void hand_off(BGrp& g, int c, uint32_t b_hi, uint32_t b_lo) {
    wait_st();
    fence_before_sync();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const uint32_t buf = g.chunk % NB;
        const uint32_t old = atom_inc_acq_rel(&g.pipe->arrivals[buf]);
        if ((old & (4u * SIB - 1u)) == 4u * SIB - 1u) {
            fence_after_sync();
            constexpr uint32_t idesc = make_idesc_tf32(128, NP);
            constexpr uint32_t LBO = 128, SBO = (KP / 4) * 128;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const uint32_t a_hi = g.tcol + buf * (4 * KC) + t * (2 * KC), a_lo = a_hi + KC;
                const uint32_t d = g.tcol + NB * 4 * KC + t * NP;
#pragma unroll
                for (int s = 0; s < KC / 8; ++s) {
                    const int ks = c * (KC / 8) + s;
                    const uint64_t bh = make_b_desc(b_hi + ks * 256, LBO, SBO);
                    const uint64_t bl = make_b_desc(b_lo + ks * 256, LBO, SBO);
                    mma_tf32_ts(d, a_lo + 8 * s, bh, idesc, ks > 0 ? 1u : 0u);      // small terms first
                    mma_tf32_ts(d, a_hi + 8 * s, bl, idesc, 1u);
                    mma_tf32_ts(d, a_hi + 8 * s, bh, idesc, 1u);
                }
            }
            mma_commit(&g.pipe->empty[buf]);
            if (c == KP / KC - 1) mma_commit(&g.pipe->dfull);
        }
    }
    __syncwarp();
    ++g.chunk;
}
template <int KC, int NB>
__device__ __forceinline__ void wait_d(BGrp& g) {
    mbar_wait(&g.pipe->dfull, g.dparity);
    g.dparity ^= 1u;
    fence_after_sync();
}
// accumulator row of this thread's pixel in tile t: NP columns (all loads in flight, one wait)
template <int NP, int KC, int NB>
__device__ __forceinline__ void load_d(const BGrp& g, int tile, float* d) {
#pragma unroll
    for (int n0 = 0; n0 < NP; n0 += 16) tmem_ld16v(g.tlane + NB * 4 * KC + tile * NP + n0, d + n0);
    wait_ld();
}

// ===================================================================================================
// s1 chain kernel
// ===================================================================================================
struct S1cArgs {
    Planes P;
    const float* w1[kMaxChain];          // tc pack of pw1 per block
    const float* w2[kMaxChain];          // tc pack of pw2
    const float* wdw[kMaxChain];         // dw3 pack
    uint32_t in_off[kMaxChain][kMaxK];   // plane offsets (floats) of branch_main's input channels
    uint32_t out_off[kMaxChain][kMaxK];  // and of its output channels
    int nblk;                            // blocks chained in this launch (> 1 only with whole-image items)
    int N, TR, bandsPerImg;
    int wbufs;                           // weight buffers in shared memory (2: next block's weights prefetched)
};

template <int K, int NP>
struct S1Smem {
    static constexpr int WFL = 2 * NP * K + 2 * NP;        // one tc pack (floats)
    static constexpr int WSET = 2 * WFL + K * 12;          // pw1 | pw2 | dw of one block
};
// T (pointwise-1 output of a band, one zero halo row above and below) is stored DENSE, even and odd rows apart:
// row tr (0 = halo above) of a plane lives in E (tr even) or O (tr odd) at [(tr/2)*W + x].  A pair's loads and stores then hit
// consecutive words in consecutive lanes (no bank conflicts); the left/right zero padding is a per-thread multiplier.
__host__ __device__ constexpr int t_half_floats(int TR, int W) { return ((TR + 3) / 2) * W + 2; }     // one of E / O, 1 pad float each side

// SIB sibling warps per TMEM lane quarter (1 or 2): warps w and w + 4 of a group share the same 32 pixel pairs; each computes KC / SIB
// of a chunk's channels and drains one of the pair's two accumulator tiles.  K = 48 ran 8 warps per SM at 232 registers (ncu:
// issue slots 37 % busy, 12 % warps active); TMEM allows no third group, siblings double the warps on the same columns.
template <int K, int NP, int G, int KC, int NB, int SIB = 1>
__global__ void __launch_bounds__(G * 128 * SIB, 1)
s1c_kernel(const __grid_constant__ S1cArgs p) {
    pdl_trigger();
    constexpr int KP = K;
    constexpr int COLS = NB * 4 * KC + 2 * NP;
    constexpr int KCS = KC / SIB, NT = G * 128 * SIB;
    static_assert(G * COLS <= 512 && K % KC == 0 && KCS % 8 == 0 && NP % 16 == 0 && K <= kMaxK && (SIB == 1 || SIB == 2), "shape");
    using L = S1Smem<K, NP>;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) BPipe pipes[G];
    __shared__ __align__(8) uint64_t wbar[2];
    __shared__ uint32_t tmem_slot;
    float* sW = smem;                                       // wbufs weight sets
    float* T = sW + (size_t)p.wbufs * L::WSET;
    const int H = p.P.H, W = p.P.W;
    const int TH = t_half_floats(p.TR, W);                  // O half starts TH floats after E
    const int TP = 2 * TH;                                  // plane stride of T
    const int warp = threadIdx.x >> 5;
    const int items = p.N * p.bandsPerImg;
    const int my_items = ((int)blockIdx.x < items) ? (items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    // weight sets this CTA consumes, in order: one per (item, block) when blocks are chained, a single one otherwise
    const int total_sets = p.nblk > 1 ? my_items * p.nblk : min(my_items, 1);

    auto load_wset = [&](int seq) {                          // one thread: weights of the seq-th (item, block) -> buffer seq % wbufs
        const int b = seq % p.nblk, buf = seq % p.wbufs;
        float* dst = sW + (size_t)buf * L::WSET;
        mbar_expect_tx(&wbar[buf], (uint32_t)(L::WSET * sizeof(float)));
        bulk_g2s(dst, p.w1[b], L::WFL * sizeof(float), &wbar[buf]);
        bulk_g2s(dst + L::WFL, p.w2[b], L::WFL * sizeof(float), &wbar[buf]);
        bulk_g2s(dst + 2 * L::WFL, p.wdw[b], K * 12 * sizeof(float), &wbar[buf]);
    };

    if (threadIdx.x == 32) {
        mbar_init(&wbar[0], 1); mbar_init(&wbar[1], 1);
        for (int i = 0; i < G; ++i) {
            mbar_init(&pipes[i].empty[0], 1); mbar_init(&pipes[i].empty[1], 1); mbar_init(&pipes[i].dfull, 1);
            pipes[i].arrivals[0] = 0; pipes[i].arrivals[1] = 0;
        }
        fence_mbar_init();
        if (total_sets > 0) load_wset(0);                   // weights do not depend on the predecessor kernel
        if (total_sets > 1 && p.wbufs > 1) load_wset(1);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    for (int i = threadIdx.x; i < K * TP; i += NT) T[i] = 0.f;      // halo rows and pads stay zero
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    BGrp g;
    const int grp = threadIdx.x / (128 * SIB);
    const int sub = SIB == 1 ? 0 : (warp >> 2) & 1;         // sibling index; the lane quarter is warp & 3 either way
    const int koff = sub * KCS;
    g.tcol = tmem_slot + grp * COLS;
    g.tlane = g.tcol + ((uint32_t)(32 * (warp & 3)) << 16);
    g.pipe = &pipes[grp];
    g.chunk = 0; g.dparity = 0;
    g.gtid = threadIdx.x & 127;
    pdl_wait();                                             // predecessor's activations are complete and visible from here on

    int seq = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int n = item / p.bandsPerImg;
        const int r0 = (item - n * p.bandsPerImg) * p.TR;
        const int rows = min(p.TR, H - r0);
        float* const img = p.P.base + (long long)n * p.P.sN + p.P.org;
        // T rows in the image: tr in [tr_lo, tr_hi]  (tr = image row - (r0 - 1))
        const int tr_lo = r0 == 0 ? 1 : 0, tr_hi = (r0 + rows == H) ? rows : rows + 1;
        if (p.bandsPerImg > 1) {
            // band items: halo rows outside the image must read as zero (a previous item may have left data there)
            if (r0 == 0) for (int i = threadIdx.x; i < K * W; i += NT) { const int k = i / W; T[k * TP + 1 + (i - k * W)] = 0.f; }
            if (r0 + rows == H) {
                const int tr = rows + 1;
                const int o = (tr & 1) * TH + 1 + (tr >> 1) * W;
                for (int i = threadIdx.x; i < K * W; i += NT) { const int k = i / W; T[k * TP + o + (i - k * W)] = 0.f; }
            }
            __syncthreads();
        }
        for (int b = 0; b < p.nblk; ++b) {
            const int wb = seq % p.wbufs;
            const float* sB1 = sW + (size_t)wb * L::WSET;
            const float* sB2 = sB1 + L::WFL;
            const float* sDW = sB2 + L::WFL;
            mbar_wait(&wbar[wb], (uint32_t)(seq / p.wbufs) & 1u);
            const uint32_t b1_hi = smem_u32(sB1), b1_lo = smem_u32(sB1 + NP * KP);
            const uint32_t b2_hi = smem_u32(sB2), b2_lo = smem_u32(sB2 + NP * KP);
            const float* aff1 = sB1 + 2 * NP * KP;              // scale[NP] | shift[NP]
            const float* aff2 = sB2 + 2 * NP * KP;
            const uint32_t* ioff = p.in_off[b];
            const uint32_t* ooff = p.out_off[b];
            // ---- phase B: pw1 + BN + ReLU on the in-image pixels of T rows [tr_lo, tr_hi] -> T ---------------------------
            // pair i <-> T rows (2i-1, 2i)  (O row i-1, E row i)
            const int i_lo = (tr_lo + 1) >> 1, i_hi = (tr_hi + 1) >> 1;
            const int nbp = (i_hi - i_lo + 1) * W;
            for (int tile = grp; tile * 128 < nbp; tile += G) {
                const int q = tile * 128 + g.gtid;
                const bool inb = q < nbp;
                const int qc = inb ? q : 0;                         // lanes past the end recompute pair 0 and drop the result
                const int ii = qc / W, x = qc - ii * W;
                const int i = i_lo + ii;
                const int tr0 = 2 * i - 1, tr1 = 2 * i;
                const bool v0 = inb && tr0 >= tr_lo, v1 = inb && tr1 <= tr_hi;
                // clamp the row of an invalid half to a valid one (its result is dropped)
                const int gr0 = r0 - 1 + (tr0 >= tr_lo ? tr0 : tr1), gr1 = r0 - 1 + (tr1 <= tr_hi ? tr1 : tr0);
                const float* ip0 = img + gr0 * p.P.Ws + x;
                const float* ip1 = img + gr1 * p.P.Ws + x;
                float v[2][2 * KCS];                                // [buffer][tile * KCS + j]: next chunk's loads fly during this chunk's hand-off
#pragma unroll
                for (int j = 0; j < KCS; ++j) { v[0][j] = __ldcg(ip0 + ioff[koff + j]); v[0][KCS + j] = __ldcg(ip1 + ioff[koff + j]); }
#pragma unroll
                for (int c = 0; c < K / KC; ++c) {
                    if (c + 1 < K / KC) {
#pragma unroll
                        for (int j = 0; j < KCS; ++j) {
                            v[(c + 1) & 1][j] = __ldcg(ip0 + ioff[(c + 1) * KC + koff + j]);
                            v[(c + 1) & 1][KCS + j] = __ldcg(ip1 + ioff[(c + 1) * KC + koff + j]);
                        }
                    }
                    acquire_buf<KC, NB>(g);
                    store_a<KC, NB, KCS>(g, 0, v[c & 1], koff);
                    store_a<KC, NB, KCS>(g, 1, v[c & 1] + KCS, koff);
                    hand_off<KP, NP, KC, NB, SIB>(g, c, b1_hi, b1_lo);
                }
                wait_d<KC, NB>(g);
                float* te = T + 1 + i * W + x;                      // E row i      (tr1)
                float* to = T + TH + 1 + (i - 1) * W + x;           // O row i - 1  (tr0)
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (SIB == 2 && t != sub) continue;             // sibling s drains tile s
                    float d[NP];
                    load_d<NP, KC, NB>(g, t, d);
                    if (t == 0 ? v0 : v1) {
                        float* tp = t == 0 ? to : te;
#pragma unroll
                        for (int n4 = 0; n4 < K; n4 += 4) {
                            const float4 sc = *reinterpret_cast<const float4*>(aff1 + n4);
                            const float4 sh = *reinterpret_cast<const float4*>(aff1 + NP + n4);
                            tp[(n4 + 0) * TP] = fmaxf(fmaf(d[n4 + 0], sc.x, sh.x), 0.f);
                            tp[(n4 + 1) * TP] = fmaxf(fmaf(d[n4 + 1], sc.y, sh.y), 0.f);
                            tp[(n4 + 2) * TP] = fmaxf(fmaf(d[n4 + 2], sc.z, sh.z), 0.f);
                            tp[(n4 + 3) * TP] = fmaxf(fmaf(d[n4 + 3], sc.w, sh.w), 0.f);
                        }
                    }
                }
            }
            __syncthreads();
            // ---- phase C: dw3x3 + BN -> pw2 + BN + ReLU -> output planes -------------------------------------------
            // pair j <-> band rows (2j, 2j+1) = T rows (2j+1, 2j+2); the stencils read T rows 2j .. 2j+3 = E[j], O[j], E[j+1], O[j+1]
            const int ncp = ((rows + 1) >> 1) * W;
            for (int tile = grp; tile * 128 < ncp; tile += G) {
                const int q = tile * 128 + g.gtid;
                const bool inb = q < ncp;
                const int qc = inb ? q : 0;
                const int j = qc / W, ox = qc - j * W;
                const bool v0 = inb, v1 = inb && 2 * j + 1 < rows;
                const float mL = ox > 0 ? 1.f : 0.f, mR = ox < W - 1 ? 1.f : 0.f;
                const float* te = T + qc;                           // E[j*W + ox - 1]  (pad float in front)
                const float* to = te + TH;
#pragma unroll 1
                for (int c = 0; c < K / KC; ++c) {
                    float a0[KCS], a1[KCS];
                    const float* e = te + (c * KC + koff) * TP;
                    const float* o = to + (c * KC + koff) * TP;
                    const float* wk = sDW + (c * KC + koff) * 12;
#pragma unroll
                    for (int jj = 0; jj < KCS; ++jj) {
                        const float4 wa = *reinterpret_cast<const float4*>(wk);
                        const float4 wb4 = *reinterpret_cast<const float4*>(wk + 4);
                        const float4 wc = *reinterpret_cast<const float4*>(wk + 8);
                        const float r0l = e[0], r0c = e[1], r0r = e[2];
                        const float r1l = o[0], r1c = o[1], r1r = o[2];
                        const float r2l = e[W], r2c = e[W + 1], r2r = e[W + 2];
                        const float r3l = o[W], r3c = o[W + 1], r3r = o[W + 2];
                        // column sums (taps w[3*dy+dx]), then the left / right columns through the border multipliers
                        float cl0 = wa.x * r0l; cl0 = fmaf(wa.w, r1l, cl0); cl0 = fmaf(wb4.z, r2l, cl0);
                        float cc0 = wa.y * r0c; cc0 = fmaf(wb4.x, r1c, cc0); cc0 = fmaf(wb4.w, r2c, cc0);
                        float cr0 = wa.z * r0r; cr0 = fmaf(wb4.y, r1r, cr0); cr0 = fmaf(wc.x, r2r, cr0);
                        float cl1 = wa.x * r1l; cl1 = fmaf(wa.w, r2l, cl1); cl1 = fmaf(wb4.z, r3l, cl1);
                        float cc1 = wa.y * r1c; cc1 = fmaf(wb4.x, r2c, cc1); cc1 = fmaf(wb4.w, r3c, cc1);
                        float cr1 = wa.z * r1r; cr1 = fmaf(wb4.y, r2r, cr1); cr1 = fmaf(wc.x, r3r, cr1);
                        const float d0 = fmaf(mR, cr0, fmaf(mL, cl0, cc0));
                        const float d1 = fmaf(mR, cr1, fmaf(mL, cl1, cc1));
                        a0[jj] = fmaf(d0, wc.y, wc.z);
                        a1[jj] = fmaf(d1, wc.y, wc.z);
                        e += TP; o += TP; wk += 12;
                    }
                    acquire_buf<KC, NB>(g);
                    store_a<KC, NB, KCS>(g, 0, a0, koff);
                    store_a<KC, NB, KCS>(g, 1, a1, koff);
                    hand_off<KP, NP, KC, NB, SIB>(g, c, b2_hi, b2_lo);
                }
                wait_d<KC, NB>(g);
                float* op = img + (r0 + 2 * j) * p.P.Ws + ox;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (SIB == 2 && t != sub) continue;
                    float d[NP];
                    load_d<NP, KC, NB>(g, t, d);
                    if (t == 0 ? v0 : v1) {
                        float* o2 = op + t * p.P.Ws;
#pragma unroll
                        for (int n4 = 0; n4 < K; n4 += 4) {
                            const float4 sc = *reinterpret_cast<const float4*>(aff2 + n4);
                            const float4 sh = *reinterpret_cast<const float4*>(aff2 + NP + n4);
                            o2[ooff[n4 + 0]] = fmaxf(fmaf(d[n4 + 0], sc.x, sh.x), 0.f);
                            o2[ooff[n4 + 1]] = fmaxf(fmaf(d[n4 + 1], sc.y, sh.y), 0.f);
                            o2[ooff[n4 + 2]] = fmaxf(fmaf(d[n4 + 2], sc.z, sh.z), 0.f);
                            o2[ooff[n4 + 3]] = fmaxf(fmaf(d[n4 + 3], sc.w, sh.w), 0.f);
                        }
                    }
                }
            }
            __syncthreads();            // T and this weight buffer are free; this block's output planes are visible to the CTA
            if (p.nblk > 1) {
                if (threadIdx.x == 32 && seq + p.wbufs < total_sets) {
                    publish_async();    // order the generic-proxy reads of the buffer before the bulk copy that overwrites it
                    load_wset(seq + p.wbufs);
                }
                ++seq;
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

// depthwise 3x3 (column stride S between taps' source pixels is 1: the taps are r[0], r[1], r[2] of three window rows) + BN
// for one channel of one output pixel; row-major staged planes.  Same association as k_tcnet.cu's dw8p (bit-identical).
template <int S>
__device__ __forceinline__ float dw3(const float* __restrict__ r0, const float* __restrict__ r1, const float* __restrict__ r2,
                                     const float* __restrict__ wk) {
    const float4 wa = *reinterpret_cast<const float4*>(wk);
    const float4 wb = *reinterpret_cast<const float4*>(wk + 4);
    const float4 wc = *reinterpret_cast<const float4*>(wk + 8);
    float p0 = wa.x * r0[0]; p0 = fmaf(wa.y, r0[1], p0); p0 = fmaf(wa.z, r0[2], p0);
    float p1 = wa.w * r1[0]; p1 = fmaf(wb.x, r1[1], p1); p1 = fmaf(wb.y, r1[2], p1);
    float p2 = wb.z * r2[0]; p2 = fmaf(wb.w, r2[1], p2); p2 = fmaf(wc.x, r2[2], p2);
    const float d = (p0 + p1) + p2;
    return fmaf(d, wc.y, wc.z);
}

// Stride-2 variant for windows whose top-left is 8-byte aligned: consecutive lanes sit two floats apart, so three scalar loads per
// window row are 2-way bank conflicted (ncu on s2c_kernel<24>: 3.2 MIO-throttle stalls per issue, 8.7 M conflicts).  The row's
// first two taps come as ONE conflict-free LDS.64 and the third is the next lane's first tap (one shuffle); lanes whose right
// neighbour is another row / another warp (`nb_ok` false) load it.  Same arithmetic and association as dw3: bit-identical.
__device__ __forceinline__ float dw3_s2_shfl(const float* __restrict__ t, int WS, const float* __restrict__ wk, bool nb_ok) {
    const float4 wa = *reinterpret_cast<const float4*>(wk);
    const float4 wb = *reinterpret_cast<const float4*>(wk + 4);
    const float4 wc = *reinterpret_cast<const float4*>(wk + 8);
    float2 v[3];
    float c2[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        v[r] = *reinterpret_cast<const float2*>(t + r * WS);
        c2[r] = __shfl_down_sync(0xffffffffu, v[r].x, 1);
        if (!nb_ok) c2[r] = t[r * WS + 2];
    }
    float p0 = wa.x * v[0].x; p0 = fmaf(wa.y, v[0].y, p0); p0 = fmaf(wa.z, c2[0], p0);
    float p1 = wa.w * v[1].x; p1 = fmaf(wb.x, v[1].y, p1); p1 = fmaf(wb.y, c2[1], p1);
    float p2 = wb.z * v[2].x; p2 = fmaf(wb.w, v[2].y, p2); p2 = fmaf(wc.x, c2[2], p2);
    const float d = (p0 + p1) + p2;
    return fmaf(d, wc.y, wc.z);
}

// ===================================================================================================
// s2c_kernel: stride-2 ShuffleV2 block (reference shufflenetv2.py:34-44,52-55), K = 24 / 48 channels per branch.
//   proj:  dw3x3 s2 + BN on the raw input -> pw + BN + ReLU             -> output planes [0, K)
//   main:  pw1 + BN + ReLU (full resolution) -> dw3x3 s2 + BN -> pw2 + BN + ReLU -> output planes [K, 2K)
// The round-1 kernel ran proj / pw1 / main as three CTA-wide phases per band with one staging buffer: ncu showed 70 % of the
// warp samples waiting (21 % on the TMA load of the band, 30 % at the phase barriers where 1 of 4 groups had a tile, 18 % on
// MMA completion).  Here a band X_i (2 TR + 1 framed input rows of the K planes, one TMA bulk copy per plane) is double
// buffered by a producer warp, pw1 runs IN PLACE in X_i, and the item loop is skewed so that both phases keep every
// warpgroup busy:
//     phase B(i):    pw1 on the in-image pixels of X_i                          (many tiles, all groups)
//     phase M+P(i):  main(i) tiles on T_i = X_i  and  proj(i+1) tiles on the raw X_(i+1), side by side
// ===================================================================================================
struct S2cArgs {
    Planes in, out;
    uint32_t in_off[kMaxK];                 // plane offsets (floats) of the K input channels
    uint32_t out_off[2 * kMaxK];            // and of the 2K output channels (proj first)
    const float* wp; const float* w1; const float* w2;      // tc packs
    const float* wdwp; const float* wdwm;                   // dw3 packs
    int N, TR, bandsPerImg;
};

template <int K, int NP, int G, int KC, int NB>
__global__ void __launch_bounds__(G * 128 + 32, 1)
s2c_kernel(const __grid_constant__ S2cArgs p) {
    pdl_trigger();
    constexpr int KP = K, NCH = K / KC;
    constexpr int COLS = NB * 2 * KC + NP, DCOL = NB * 2 * KC;
    static_assert(G * COLS <= 512 && K % KC == 0 && KC % 8 == 0 && NP % 16 == 0 && K <= kMaxK, "shape");
    constexpr int WFL = 2 * NP * KP + 2 * NP;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) BPipe pipes[G];
    __shared__ __align__(8) uint64_t wbar, xfull[2], xfree[2];
    __shared__ uint32_t tmem_slot;
    float* sBp = smem;
    float* sB1 = sBp + WFL;
    float* sB2 = sB1 + WFL;
    float* sDWp = sB2 + WFL;
    float* sDWm = sDWp + K * 12;
    float* Xb = sDWm + K * 12;
    const int Hin = p.in.H, Win = p.in.W, WS = p.in.Ws, pin = p.in.pad;
    const int Hout = p.out.H, Wout = p.out.W;
    const int XR = 2 * p.TR + 1;                            // staged rows per band
    const int RS = XR * WS;                                 // plane stride inside a staging buffer
    const int XBUF = K * RS;
    const int warp = threadIdx.x >> 5;
    const int items = p.N * p.bandsPerImg;
    const int my_items = ((int)blockIdx.x < items) ? (items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        mbar_init(&wbar, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xfree[i], G * 4); }
        for (int i = 0; i < G; ++i) {
            mbar_init(&pipes[i].empty[0], 1); mbar_init(&pipes[i].empty[1], 1); mbar_init(&pipes[i].dfull, 1);
            pipes[i].arrivals[0] = 0; pipes[i].arrivals[1] = 0;
        }
        fence_mbar_init();
        mbar_expect_tx(&wbar, (uint32_t)((3 * WFL + 2 * K * 12) * sizeof(float)));
        bulk_g2s(sBp, p.wp, WFL * sizeof(float), &wbar);
        bulk_g2s(sB1, p.w1, WFL * sizeof(float), &wbar);
        bulk_g2s(sB2, p.w2, WFL * sizeof(float), &wbar);
        bulk_g2s(sDWp, p.wdwp, K * 12 * sizeof(float), &wbar);
        bulk_g2s(sDWm, p.wdwm, K * 12 * sizeof(float), &wbar);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    pdl_wait();                                             // predecessor's activations are complete and visible from here on

    auto item_geom = [&](int it, int& n, int& r0, int& rows) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        n = item / p.bandsPerImg;
        r0 = (item - n * p.bandsPerImg) * p.TR;
        rows = min(p.TR, Hout - r0);
    };

    if (warp == G * 4) {
        // ---------------- producer: band it -> buffer it & 1 -------------------------------------------------------
        const int lane = threadIdx.x & 31;
        for (int it = 0; it < my_items; ++it) {
            const int buf = it & 1;
            int n, r0, rows;
            item_geom(it, n, r0, rows);
            const int nrows = 2 * rows + 1;
            if (it >= 2) mbar_wait(&xfree[buf], (uint32_t)((it >> 1) - 1) & 1u);     // every compute warp is done with the buffer
            publish_async();
            if (lane == 0) mbar_expect_tx(&xfull[buf], (uint32_t)(K * nrows * WS * sizeof(float)));
            __syncwarp();
            const float* src0 = p.in.base + (long long)n * p.in.sN + (long long)(2 * r0 - 1 + pin) * WS;
            for (int k = lane; k < K; k += 32)
                bulk_g2s(Xb + (size_t)buf * XBUF + (size_t)k * RS, src0 + p.in_off[k], (uint32_t)(nrows * WS * sizeof(float)), &xfull[buf]);
        }
    } else {
        // ---------------- compute warpgroups -----------------------------------------------------------------------
        BGrp g;
        const int grp = threadIdx.x >> 7;
        g.tcol = tmem_slot + grp * COLS;
        g.tlane = g.tcol + ((uint32_t)(32 * (warp & 3)) << 16);
        g.pipe = &pipes[grp];
        g.chunk = 0; g.dparity = 0;
        g.gtid = threadIdx.x & 127;
        const uint32_t bp_hi = smem_u32(sBp), bp_lo = smem_u32(sBp + NP * KP);
        const uint32_t b1_hi = smem_u32(sB1), b1_lo = smem_u32(sB1 + NP * KP);
        const uint32_t b2_hi = smem_u32(sB2), b2_lo = smem_u32(sB2 + NP * KP);
        const float* affp = sBp + 2 * NP * KP;
        const float* aff1 = sB1 + 2 * NP * KP;
        const float* aff2 = sB2 + 2 * NP * KP;
        mbar_wait(&wbar, 0);

        // one 128-pixel tile of a depthwise-s2 -> pointwise branch of band (n, r0, rows) out of staging buffer X
        auto dw_tile = [&](const float* X, int n, int r0, int rows, int tile, const float* sDW, uint32_t b_hi, uint32_t b_lo,
                           const float* aff, const uint32_t* ooff) {
            const int npix = rows * Wout;
            const int q = tile * 128 + g.gtid;
            const bool valid = q < npix;
            const int qc = valid ? q : 0;
            const int orow = qc / Wout, ox = qc - orow * Wout;
            const float* w0 = X + (2 * orow) * WS + 2 * ox + (pin - 1);      // window's top-left (frame column pin-1 <-> input column -1)
            const bool aligned = (pin & 1) != 0;                               // frame of 1: every window starts on an even float
            const bool nb_ok = ox + 1 < Wout && (threadIdx.x & 31) != 31;      // the next lane holds the pixel to the right
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                float a[KC];
                const float* t = w0 + c * KC * RS;
                const float* wk = sDW + c * KC * 12;
                if (aligned) {
#pragma unroll
                    for (int j = 0; j < KC; ++j) { a[j] = dw3_s2_shfl(t, WS, wk, nb_ok); t += RS; wk += 12; }
                } else {
#pragma unroll
                    for (int j = 0; j < KC; ++j) { a[j] = dw3<2>(t, t + WS, t + 2 * WS, wk); t += RS; wk += 12; }
                }
                st_acquire<NB>(g);
                st_store<KC, NB>(g, a);
                st_hand_off<KP, NP, KC, NB>(g, c, b_hi, b_lo, DCOL, c == NCH - 1);
            }
            st_wait_d(g);
            float* op = p.out.base + (long long)n * p.out.sN + p.out.org + (r0 + orow) * p.out.Ws + ox;
#pragma unroll
            for (int n0 = 0; n0 < NP; n0 += 16) {
                if (n0 >= K) break;
                float d[16];
                tmem_ld16v(g.tlane + DCOL + n0, d);
                wait_ld();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + j < K) op[ooff[n0 + j]] = fmaxf(fmaf(d[j], aff[n0 + j], aff[NP + n0 + j]), 0.f);
                }
            }
        };
        auto proj_tiles = [&](int it, int first, int step) {          // proj tiles first, first+step, ... of band it
            int n, r0, rows;
            item_geom(it, n, r0, rows);
            const float* X = Xb + (size_t)(it & 1) * XBUF;
            for (int tile = first; tile * 128 < rows * Wout; tile += step)
                dw_tile(X, n, r0, rows, tile, sDWp, bp_hi, bp_lo, affp, p.out_off);
        };

        if (my_items > 0) {
            mbar_wait(&xfull[0], 0);
            proj_tiles(0, grp, G);
            group_bar(1, G * 128);                          // proj(0) has read the raw X_0: pw1 may overwrite it
        }
        for (int it = 0; it < my_items; ++it) {
            int n, r0, rows;
            item_geom(it, n, r0, rows);
            float* X = Xb + (size_t)(it & 1) * XBUF;
            const int gr0 = 2 * r0 - 1, nrows = 2 * rows + 1;     // input rows [gr0, gr0 + nrows) <-> staged rows [0, nrows)
            // ---- phase B: pw1 + BN + ReLU in place on every staged in-image pixel (frame / halo positions stay zero) --------
            const int gr_lo = max(gr0, 0), gr_hi = min(gr0 + nrows - 1, Hin - 1);
            const int npos = (gr_hi - gr_lo + 1) * Win;
            for (int tile = grp; tile * 128 < npos; tile += G) {
                const int q = tile * 128 + g.gtid;
                const bool valid = q < npos;
                const int qc = valid ? q : 0;
                const int rr = qc / Win, x = qc - rr * Win;
                float* tpos = X + (gr_lo + rr - gr0) * WS + pin + x;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    float a[KC];
#pragma unroll
                    for (int j = 0; j < KC; ++j) a[j] = tpos[(c * KC + j) * RS];
                    st_acquire<NB>(g);
                    st_store<KC, NB>(g, a);
                    st_hand_off<KP, NP, KC, NB>(g, c, b1_hi, b1_lo, DCOL, c == NCH - 1);
                }
                st_wait_d(g);
#pragma unroll
                for (int n0 = 0; n0 < NP; n0 += 16) {
                    if (n0 >= K) break;
                    float d[16];
                    tmem_ld16v(g.tlane + DCOL + n0, d);
                    wait_ld();
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + j < K) tpos[(n0 + j) * RS] = fmaxf(fmaf(d[j], aff1[n0 + j], aff1[NP + n0 + j]), 0.f);
                    }
                }
            }
            group_bar(1, G * 128);
            // ---- phase M+P: main(it) out of T = X, proj(it + 1) out of the raw next band, side by side ---------------------
            const int mt = (rows * Wout + 127) / 128;               // main tiles of this band
            const bool more = it + 1 < my_items;
            if (more) mbar_wait(&xfull[(it + 1) & 1], (uint32_t)((it + 1) >> 1) & 1u);
            for (int job = grp; job < mt; job += G)
                dw_tile(X, n, r0, rows, job, sDWm, b2_hi, b2_lo, aff2, p.out_off + K);
            if (more) proj_tiles(it + 1, ((grp - mt) % G + G) % G, G);       // the groups after the main jobs take the first proj tiles
            group_bar(1, G * 128);                          // X_it is fully consumed; proj(it + 1) has read the raw X_(it+1)
            if ((threadIdx.x & 31) == 0) mbar_arrive(&xfree[it & 1]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

// ===================================================================================================
// pw3_kernel: plain pointwise convolution over planes, out = ReLU(BN(W . in)), with gather-on-load: KA channels read through
// (y >> SHA, x >> SHA) from pool A (the FPN's nearest-neighbour up-sampling of C3, reference model/fpn.py:57-58), then KB channels
// at (y, x) from pool B.  Used by fpn.S2 (192 up-sampled + 96), fpn.S3 (192) and the pointwise-1 of the K = 96 stride-2 block.
// One tile = 128 consecutive pixels of the flattened (image, pixel) space; a thread's K loads are software pipelined, the
// next 16 channels are in flight while the current 16 are split and handed to the tensor core (round 1 issued 48 loads,
// waited, then paid 6 hand-offs of 8 channels: six fully exposed load batches per 288-channel tile).
// ===================================================================================================
constexpr int kPwMaxK = 288;
struct Pw3Args {
    Planes A, B, out;
    uint32_t in_off[kPwMaxK];          // plane offset (floats) of input channel k inside its pool
    uint32_t out_off[96];
    const float* wpack;                // tc pack: Bhi | Blo | scale | shift
    int N, nout;
};

// SIB = 2: sibling warps as in s1c_kernel.  TMEM (G column blocks of NB*2*KC + NP) caps the kernel at three warpgroups = 12 warps
// per SM, and ncu (profiles/r2_final_kernels_ncu.txt) shows them waiting on their global loads (long scoreboard 2.4-3.1 per issue,
// issue slots 36-38 % busy).  Warps w and w + 4 of a group share a lane quarter: each loads, splits and stores half of a chunk's
// channels, the last of the EIGHT warps issues the chunk's MMAs, and each drains half of the accumulator columns.
template <int KA, int KB, int SHA, int NP, int G, bool RELU, int KC = 16, int SIB = 1>
__global__ void __launch_bounds__(G * 128 * SIB, 1)
pw3_kernel(const __grid_constant__ Pw3Args p) {
    pdl_trigger();
    constexpr int KP = KA + KB, NB = 2, NCH = KP / KC, KCS = KC / SIB;
    constexpr int COLS = NB * 2 * KC + NP;
    static_assert(G * COLS <= 512 && KP % (2 * KC) == 0 && KA % KC == 0 && NP % 16 == 0 && KP <= kPwMaxK && KCS % 8 == 0 && (SIB == 1 || SIB == 2), "shape");
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) BPipe pipes[G];
    __shared__ __align__(8) uint64_t wbar;
    __shared__ uint32_t tmem_slot;
    float* sB = smem;
    constexpr int WFL = 2 * NP * KP + 2 * NP;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 32) {
        mbar_init(&wbar, 1);
        for (int i = 0; i < G; ++i) {
            mbar_init(&pipes[i].empty[0], 1); mbar_init(&pipes[i].empty[1], 1); mbar_init(&pipes[i].dfull, 1);
            pipes[i].arrivals[0] = 0; pipes[i].arrivals[1] = 0;
        }
        fence_mbar_init();
        mbar_expect_tx(&wbar, (uint32_t)(WFL * sizeof(float)));
        constexpr uint32_t kPiece = 32768;                 // a bulk copy carries at most 2^20-1 bytes; keep the pieces modest
        for (uint32_t o = 0; o < WFL * sizeof(float); o += kPiece)
            bulk_g2s(reinterpret_cast<char*>(sB) + o, reinterpret_cast<const char*>(p.wpack) + o,
                     min(kPiece, (uint32_t)(WFL * sizeof(float)) - o), &wbar);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    BGrp g;
    const int grp = threadIdx.x / (128 * SIB);
    const int sub = SIB == 1 ? 0 : (warp >> 2) & 1;        // sibling index; the lane quarter is warp & 3 either way
    const int koff = sub * KCS;
    g.tcol = tmem_slot + grp * COLS;
    g.tlane = g.tcol + ((uint32_t)(32 * (warp & 3)) << 16);
    g.pipe = &pipes[grp];
    g.chunk = 0; g.dparity = 0;
    g.gtid = threadIdx.x & 127;
    const uint32_t b_hi = smem_u32(sB), b_lo = smem_u32(sB + NP * KP);
    const float* aff = sB + 2 * NP * KP;
    const int HW = p.out.H * p.out.W, W = p.out.W;
    const long long total = (long long)p.N * HW;
    const int ntiles = (int)((total + 127) / 128);
    // accumulator columns this warp drains: everything, or the sibling's half (in blocks of 16)
    constexpr int NBLK = NP / 16, H0 = (NBLK + 1) / 2;
    const int n_lo = SIB == 1 ? 0 : (sub ? H0 * 16 : 0), n_hi = SIB == 1 ? NP : (sub ? NP : H0 * 16);
    pdl_wait();                                            // predecessor's activations are complete and visible from here on
    mbar_wait(&wbar, 0);
    for (int tile = blockIdx.x * G + grp; tile < ntiles; tile += gridDim.x * G) {
        const long long pos = (long long)tile * 128 + g.gtid;
        const bool valid = pos < total;
        const long long pc = valid ? pos : 0;              // lanes past the end recompute pixel 0 and drop the result
        const int n = (int)(pc / HW);
        const int px = (int)(pc - (long long)n * HW);
        const int y = px / W, x = px - y * W;
        const float* baseA = p.A.base + (long long)n * p.A.sN + p.A.org + (y >> SHA) * p.A.Ws + (x >> SHA);
        const float* baseB = p.B.base + (long long)n * p.B.sN + p.B.org + y * p.B.Ws + x;
        float v[2][KCS];
        auto load = [&](int c, float* dst) {               // c is a multiple-of-KC chunk index; chunks never straddle the A / B split
            const float* base = (c * KC < KA) ? baseA : baseB;
#pragma unroll
            for (int j = 0; j < KCS; ++j) dst[j] = __ldcg(base + p.in_off[c * KC + koff + j]);
        };
        auto store = [&](const float* a) {                  // this warp's KCS channels of the chunk -> hi / lo columns of the current A buffer
            const uint32_t col = g.tlane + (g.chunk % NB) * (2 * KC) + koff;
#pragma unroll
            for (int j = 0; j < KCS; j += 8) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    hi[i] = __float_as_uint(a[j + i]) & 0xFFFFE000u;
                    lo[i] = __float_as_uint(a[j + i] - __uint_as_float(hi[i]));
                }
                tmem_st8(col + j, hi);
                tmem_st8(col + KC + j, lo);
            }
        };
        auto hand_off = [&](int c, bool last) {            // st_hand_off with 4 * SIB arriving warps
            wait_st();
            fence_before_sync();
            __syncwarp();
            if ((threadIdx.x & 31) == 0) {
                const uint32_t buf = g.chunk % NB;
                const uint32_t old = atom_inc_acq_rel(&g.pipe->arrivals[buf]);
                if ((old & (4u * SIB - 1u)) == 4u * SIB - 1u) {
                    fence_after_sync();
                    constexpr uint32_t idesc = make_idesc_tf32(128, NP);
                    constexpr uint32_t LBO = 128, SBO = (KP / 4) * 128;
                    const uint32_t a_hi = g.tcol + buf * (2 * KC), a_lo = a_hi + KC, d = g.tcol + NB * 2 * KC;
#pragma unroll
                    for (int s2 = 0; s2 < KC / 8; ++s2) {
                        const int ks = c * (KC / 8) + s2;
                        const uint64_t bh = make_b_desc(b_hi + ks * 256, LBO, SBO);
                        const uint64_t bl = make_b_desc(b_lo + ks * 256, LBO, SBO);
                        mma_tf32_ts(d, a_lo + 8 * s2, bh, idesc, ks > 0 ? 1u : 0u);      // small terms first
                        mma_tf32_ts(d, a_hi + 8 * s2, bl, idesc, 1u);
                        mma_tf32_ts(d, a_hi + 8 * s2, bh, idesc, 1u);
                    }
                    mma_commit(&g.pipe->empty[buf]);
                    if (last) mma_commit(&g.pipe->dfull);
                }
            }
            __syncwarp();
            ++g.chunk;
        };
        load(0, v[0]);
#pragma unroll 1
        for (int c = 0; c < NCH; c += 2) {
            load(c + 1, v[1]);
            st_acquire<NB>(g);
            store(v[0]);
            hand_off(c, false);
            if (c + 2 < NCH) load(c + 2, v[0]);
            st_acquire<NB>(g);
            store(v[1]);
            hand_off(c + 1, c + 2 == NCH);
        }
        st_wait_d(g);
        float* obase = p.out.base + (long long)n * p.out.sN + p.out.org + y * p.out.Ws + x;
#pragma unroll
        for (int n0 = 0; n0 < NP; n0 += 16) {
            if (n0 < n_lo || n0 >= n_hi) continue;
            float d[16];
            tmem_ld16v(g.tlane + NB * 2 * KC + n0, d);
            wait_ld();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int nn = n0 + j;
                    if (nn < p.nout) {
                        float r = fmaf(d[j], aff[nn], aff[NP + nn]);
                        if (RELU) r = fmaxf(r, 0.f);
                        obase[p.out_off[nn]] = r;
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

template <typename Kern>
int blk_smem_attr(Kern kern, size_t bytes) {
    if (bytes > kSmemCap) { set_error("block kernel needs %zu bytes of shared memory", bytes); return YFV2_EUNSUPPORTED; }
    YFV2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return YFV2_OK;
}

constexpr size_t kBlkSmemBudget = kSmemCap - 1024;          // static shared memory (pipes, barriers) + slack

// Shared-memory plan of the stride-1 kernel for an H x W map: whole images (chainable) when they fit.
struct S1Geom { int TR, bands, wbufs; size_t bytes; bool whole; };
S1Geom s1_geometry(int K, int H, int W, int nblk) {
    const int NP = tc_round(K, 16);
    const size_t wset = (size_t)2 * (2 * NP * K + 2 * NP) + K * 12;
    auto need = [&](int tr, int wbufs) { return (wbufs * wset + (size_t)K * 2 * t_half_floats(tr, W)) * sizeof(float); };
    S1Geom g{};
    if (need(H, 1) <= kBlkSmemBudget) {
        g.whole = true; g.TR = H; g.bands = 1;
        g.wbufs = (nblk > 1 && need(H, 2) <= kBlkSmemBudget) ? 2 : 1;
        g.bytes = need(H, g.wbufs);
        return g;
    }
    g.whole = false; g.wbufs = 1;
    for (int nb = 2; nb <= H; ++nb) {
        const int tr = ((H + nb - 1) / nb + 1) & ~1;            // even: a pair never straddles two bands
        if (need(tr, 1) <= kBlkSmemBudget) { g.TR = tr; g.bands = (H + tr - 1) / tr; g.bytes = need(tr, 1); return g; }
    }
    g.TR = 0;
    return g;
}

}  // namespace

// true if consecutive stride-1 blocks of width K on an H x W map can share one launch
bool blk_s1_chainable(int K, int H, int W) { return K <= kMaxK && s1_geometry(K, H, W, 2).whole; }

#define TRYB(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

// Runs up to nblk consecutive stride-1 blocks of branch width K (24 or 48) over pool P.  *done = blocks actually fused
// into this launch (all of them when an image's T fits in shared memory, otherwise 1).
int blk_launch_s1(int K, const Planes& P, int nblk, const ChanTab* tin, const ChanTab* tout, const float* const* w1,
                  const float* const* wdw, const float* const* w2, int N, cudaStream_t s, int* done) {
    if (K != 24 && K != 48) { set_error("blk_launch_s1: unsupported K=%d", K); return YFV2_EUNSUPPORTED; }
    if (nblk > kMaxChain) nblk = kMaxChain;
    S1Geom geo = s1_geometry(K, P.H, P.W, nblk);
    if (geo.TR <= 0) { set_error("blk_launch_s1: a %dx%d map does not fit in shared memory", P.H, P.W); return YFV2_EUNSUPPORTED; }
    if (!geo.whole) nblk = 1;
    S1cArgs a{};
    a.P = P; a.nblk = nblk; a.N = N; a.TR = geo.TR; a.bandsPerImg = geo.bands; a.wbufs = geo.wbufs;
    for (int b = 0; b < nblk; ++b) {
        a.w1[b] = w1[b]; a.w2[b] = w2[b]; a.wdw[b] = wdw[b];
        for (int k = 0; k < K; ++k) {
            a.in_off[b][k] = (uint32_t)((long long)tin[b].c[k] * P.sC);
            a.out_off[b][k] = (uint32_t)((long long)tout[b].c[k] * P.sC);
        }
    }
    const int items = N * geo.bands;
    auto run = [&](auto kern, int threads) -> int {
        TRYB(blk_smem_attr(kern, geo.bytes));
        YFV2_CUDA(launch_k(kern, min(items, sm_count()), threads, geo.bytes, s, pdl_take(), a));
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    *done = nblk;
    static const bool no_sib = getenv("YFV2_S1_NOSIB") != nullptr;          // one warp per lane quarter, kept for A/B runs
    if (K == 24) return run(s1c_kernel<24, 32, 4, 8, 2>, 4 * 128);
    if (no_sib) return run(s1c_kernel<48, 48, 2, 16, 2>, 2 * 128);
    return run(s1c_kernel<48, 48, 2, 16, 2, 2>, 2 * 256);
}

// Stride-2 block of branch width K (24 or 48): in (K planes at 2H x 2W) -> out (2K planes at H x W).
int blk_launch_s2(int K, const Planes& in, const Planes& out, const ChanTab& tin, const ChanTab& tout, const float* wdwp, const float* wp,
                  const float* w1, const float* wdwm, const float* w2, int N, cudaStream_t s) {
    if (K != 24 && K != 48) { set_error("blk_launch_s2: unsupported K=%d", K); return YFV2_EUNSUPPORTED; }
    S2cArgs a{};
    a.in = in; a.out = out; a.wp = wp; a.w1 = w1; a.w2 = w2; a.wdwp = wdwp; a.wdwm = wdwm; a.N = N;
    for (int k = 0; k < K; ++k) a.in_off[k] = (uint32_t)((long long)tin.c[k] * in.sC);
    for (int k = 0; k < 2 * K; ++k) a.out_off[k] = (uint32_t)((long long)tout.c[k] * out.sC);
    const int NP = tc_round(K, 16);
    const size_t wfl = (size_t)3 * (2 * NP * K + 2 * NP) + 2 * K * 12;
    auto bytes = [&](int tr) { return (wfl + (size_t)2 * K * (2 * tr + 1) * in.Ws) * sizeof(float); };
    int TR = out.H;
    while (TR > 1 && bytes(TR) > kBlkSmemBudget) --TR;
    if (bytes(TR) > kBlkSmemBudget) { set_error("blk_launch_s2: a %dx%d input does not fit in shared memory", in.H, in.W); return YFV2_EUNSUPPORTED; }
    const int bands = (out.H + TR - 1) / TR;
    TR = (out.H + bands - 1) / bands;                        // equalise the bands
    a.TR = TR; a.bandsPerImg = (out.H + TR - 1) / TR;
    const int items = N * a.bandsPerImg;
    auto run = [&](auto kern, int G) -> int {
        TRYB(blk_smem_attr(kern, bytes(TR)));
        YFV2_CUDA(launch_k(kern, min(items, sm_count()), G * 128 + 32, bytes(TR), s, pdl_take(), a));
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    if (K == 24) return run(s2c_kernel<24, 32, 4, 24, 1>, 4);
    return run(s2c_kernel<48, 48, 4, 16, 2>, 4);
}

// plain pointwise; kind 0: 96->96 (+ReLU)  1: FPN S3 192->72 (+ReLU)  2: FPN S2 (up(192) ++ 96)->72 (+ReLU)  3: 48->48 (+ReLU)
int blk_launch_pw(int kind, const Planes& A, const ChanTab& ta, const Planes& B, const ChanTab& tb, const Planes& out, const ChanTab& tout,
                  const float* wpack, int N, cudaStream_t s) {
    Pw3Args a{};
    a.A = A; a.B = B; a.out = out; a.wpack = wpack; a.N = N;
    const long long total = (long long)N * out.H * out.W;
    const int ntiles = (int)((total + 127) / 128);
    auto run = [&](auto kern, int KA, int KB, int NP, int G, int nout, int sib = 1) -> int {
        a.nout = nout;
        for (int k = 0; k < KA; ++k) a.in_off[k] = (uint32_t)((long long)ta.c[k] * A.sC);
        for (int k = 0; k < KB; ++k) a.in_off[KA + k] = (uint32_t)((long long)tb.c[k] * B.sC);
        for (int k = 0; k < nout; ++k) a.out_off[k] = (uint32_t)((long long)tout.c[k] * out.sC);
        const size_t bytes = (size_t)(2 * NP * (KA + KB) + 2 * NP) * sizeof(float);
        TRYB(blk_smem_attr(kern, bytes));
        YFV2_CUDA(launch_k(kern, min((ntiles + G - 1) / G, sm_count()), G * 128 * sib, bytes, s, pdl_take(), a));
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    // sibling warps on three warpgroups (24 warps, 76-79 registers) by default: stage4.0 109.9 -> 105.9 us, fpn.S3 26.8 -> 25.2,
    // fpn.S2 73.8 -> 69.7 at batch 256.  YFV2_PW_SIB=0: none (12 warps), 2: siblings on two warpgroups (16 warps; fpn.S2 78.5 us)
    static const int pw_sib = getenv("YFV2_PW_SIB") ? atoi(getenv("YFV2_PW_SIB")) : 1;
    if (pw_sib == 1) {
        if (kind == 0) return run(pw3_kernel<96, 0, 0, 96, 3, true, 16, 2>, 96, 0, 96, 3, 96, 2);
        if (kind == 1) return run(pw3_kernel<192, 0, 0, 80, 3, true, 16, 2>, 192, 0, 80, 3, 72, 2);
        if (kind == 2) return run(pw3_kernel<192, 96, 1, 80, 3, true, 16, 2>, 192, 96, 80, 3, 72, 2);
    } else if (pw_sib == 2) {
        if (kind == 0) return run(pw3_kernel<96, 0, 0, 96, 2, true, 16, 2>, 96, 0, 96, 2, 96, 2);
        if (kind == 1) return run(pw3_kernel<192, 0, 0, 80, 2, true, 16, 2>, 192, 0, 80, 2, 72, 2);
        if (kind == 2) return run(pw3_kernel<192, 96, 1, 80, 2, true, 16, 2>, 192, 96, 80, 2, 72, 2);
    }
    if (kind == 0) return run(pw3_kernel<96, 0, 0, 96, 3, true>, 96, 0, 96, 3, 96);
    if (kind == 1) return run(pw3_kernel<192, 0, 0, 80, 3, true>, 192, 0, 80, 3, 72);
    if (kind == 2) return run(pw3_kernel<192, 96, 1, 80, 3, true>, 192, 96, 80, 3, 72);
    if (kind == 3) return run(pw3_kernel<48, 0, 0, 48, 4, true, 8>, 48, 0, 48, 4, 48);      // pw1 of the K=48 stride-2 block
    set_error("blk_launch_pw: unknown kind %d", kind);
    return YFV2_EUNSUPPORTED;
}

}  // namespace yfv2
