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
// KC channel values of one of this thread's two pixels -> tf32 hi / lo columns of the current A buffer.
template <int KC, int NB>
__device__ __forceinline__ void store_a(const BGrp& g, int tile, const float* a) {
    const uint32_t col = g.tlane + (g.chunk % NB) * (4 * KC) + tile * (2 * KC);
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
// Chunk c (of KP / KC) of both tiles is in TMEM: the last of the group's four warps to get here issues the MMAs.
template <int KP, int NP, int KC, int NB>
__device__ __forceinline__ void hand_off(BGrp& g, int c, uint32_t b_hi, uint32_t b_lo) {
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

template <int K, int NP, int G, int KC, int NB>
__global__ void __launch_bounds__(G * 128, 1)
s1c_kernel(const __grid_constant__ S1cArgs p) {
    pdl_trigger();
    constexpr int KP = K;
    constexpr int COLS = NB * 4 * KC + 2 * NP;
    static_assert(G * COLS <= 512 && K % KC == 0 && KC % 8 == 0 && NP % 16 == 0 && K <= kMaxK, "shape");
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
    for (int i = threadIdx.x; i < K * TP; i += G * 128) T[i] = 0.f;      // halo rows and pads stay zero
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    BGrp g;
    const int grp = threadIdx.x >> 7;
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
            if (r0 == 0) for (int i = threadIdx.x; i < K * W; i += G * 128) { const int k = i / W; T[k * TP + 1 + (i - k * W)] = 0.f; }
            if (r0 + rows == H) {
                const int tr = rows + 1;
                const int o = (tr & 1) * TH + 1 + (tr >> 1) * W;
                for (int i = threadIdx.x; i < K * W; i += G * 128) { const int k = i / W; T[k * TP + o + (i - k * W)] = 0.f; }
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
                float v[2][2 * KC];                                 // [buffer][tile * KC + j]: next chunk's loads fly during this chunk's hand-off
#pragma unroll
                for (int j = 0; j < KC; ++j) { v[0][j] = __ldcg(ip0 + ioff[j]); v[0][KC + j] = __ldcg(ip1 + ioff[j]); }
#pragma unroll
                for (int c = 0; c < K / KC; ++c) {
                    if (c + 1 < K / KC) {
#pragma unroll
                        for (int j = 0; j < KC; ++j) {
                            v[(c + 1) & 1][j] = __ldcg(ip0 + ioff[(c + 1) * KC + j]);
                            v[(c + 1) & 1][KC + j] = __ldcg(ip1 + ioff[(c + 1) * KC + j]);
                        }
                    }
                    acquire_buf<KC, NB>(g);
                    store_a<KC, NB>(g, 0, v[c & 1]);
                    store_a<KC, NB>(g, 1, v[c & 1] + KC);
                    hand_off<KP, NP, KC, NB>(g, c, b1_hi, b1_lo);
                }
                wait_d<KC, NB>(g);
                float* te = T + 1 + i * W + x;                      // E row i      (tr1)
                float* to = T + TH + 1 + (i - 1) * W + x;           // O row i - 1  (tr0)
#pragma unroll
                for (int t = 0; t < 2; ++t) {
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
                    float a0[KC], a1[KC];
                    const float* e = te + c * KC * TP;
                    const float* o = to + c * KC * TP;
                    const float* wk = sDW + c * KC * 12;
#pragma unroll
                    for (int jj = 0; jj < KC; ++jj) {
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
                    store_a<KC, NB>(g, 0, a0);
                    store_a<KC, NB>(g, 1, a1);
                    hand_off<KP, NP, KC, NB>(g, c, b2_hi, b2_lo);
                }
                wait_d<KC, NB>(g);
                float* op = img + (r0 + 2 * j) * p.P.Ws + ox;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
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
    auto run = [&](auto kern, int G) -> int {
        TRYB(blk_smem_attr(kern, geo.bytes));
        YFV2_CUDA(launch_k(kern, min(items, sm_count()), G * 128, geo.bytes, s, pdl_take(), a));
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    *done = nblk;
    if (K == 24) return run(s1c_kernel<24, 32, 4, 8, 2>, 4);
    return run(s1c_kernel<48, 48, 2, 16, 2>, 2);
}

}  // namespace yfv2
