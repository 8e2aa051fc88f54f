// Small-map chains: the stride-1 ShuffleV2 blocks of stage 4 (K = 96 per branch) on maps of at most 128 pixels
// (11x11 for a 352x352 input; reference model/backbone/shufflenetv2.py:19-32,48-51,82-95).
//
// Why a kernel of its own: at this size the weights (two 96x96 tf32 hi/lo packs = 149 KB) dwarf the activations
// (96 x 121 floats per image and branch), and one image is ONE 128-row MMA tile.  Round 1 ran every block as two
// launches (pw1 to scratch planes, then dw->pw2) of kernels built for large maps: 80 us per block for 48 MB of
// traffic (0.09 of the HBM roofline).  Here a warpgroup owns a whole image (thread <-> pixel <-> TMEM lane), so
// the depthwise stencil's neighbours all live in the same warpgroup and nothing but a 128-thread named barrier is
// needed between the pointwise-1 epilogue and the stencil:
//   phase B  pw1: 96 input channels of my pixel (L2) -> split -> TMEM ring (16-channel chunks) -> D1[96]
//   phase C  per 16-channel chunk: D1 columns -> BN+ReLU -> shared T chunk (dense rows, zero halo rows; double
//            buffered) -> named barrier -> dw3x3+BN out of T -> TMEM ring -> pw2 accumulates that K-chunk into D2[96]
//            D2 -> BN+ReLU -> output planes
// Two warpgroups (two images) share one CTA and one copy of the weights; a block's output pixel is written and, in the
// next block of the chain, read by the same thread, so the blocks of a stage chain inside one launch without any
// grid- or CTA-wide synchronisation of activations.  A ninth warp streams the next block's weights in as soon as
// both groups have finished with a pack (pw1 pack: after phase B; pw2 + dw packs: after phase C).
#include "eng3.cuh"

namespace yfv2 {
namespace {

using namespace tc;
using namespace eng3;

constexpr int kTailMaxBlocks = 4;
constexpr int kTK = 96;                 // branch width
constexpr int kTKC = 16;                // channels per hand-off
constexpr int kTRing = 2 * 2 * kTKC;    // 2 buffers x (hi + lo)
constexpr int kTCols = kTRing + 2 * kTK;      // ring | D1 | D2 = 256 columns per group

struct TailArgs {
    Planes P;
    const float* w1[kTailMaxBlocks];
    const float* w2[kTailMaxBlocks];
    const float* wdw[kTailMaxBlocks];
    uint32_t in_off[kTailMaxBlocks][kTK];
    uint32_t out_off[kTailMaxBlocks][kTK];
    int nblk, N;
};

constexpr int kTWFL = 2 * kTK * kTK + 2 * kTK;          // one tc pack (floats)

// ring hand-off of one 16-channel chunk into accumulator column dcol (see k_blk.cu hand_off)
__device__ __forceinline__ void tail_store(const BGrp& g, const float* a) {
    const uint32_t col = g.tlane + (g.chunk & 1u) * (2 * kTKC);
#pragma unroll
    for (int j = 0; j < kTKC; j += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            hi[i] = __float_as_uint(a[j + i]) & 0xFFFFE000u;
            lo[i] = __float_as_uint(a[j + i] - __uint_as_float(hi[i]));
        }
        tmem_st8(col + j, hi);
        tmem_st8(col + kTKC + j, lo);
    }
}
__device__ __forceinline__ void tail_acquire(BGrp& g) {
    const uint32_t buf = g.chunk & 1u, use = g.chunk >> 1;
    if (use > 0) mbar_wait(&g.pipe->empty[buf], (use - 1) & 1u);
    fence_after_sync();
}
__device__ __forceinline__ void tail_hand_off(BGrp& g, int c, uint32_t b_hi, uint32_t b_lo, uint32_t dcol) {
    wait_st();
    fence_before_sync();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const uint32_t buf = g.chunk & 1u;
        const uint32_t old = atom_inc_acq_rel(&g.pipe->arrivals[buf]);
        if ((old & 3u) == 3u) {
            fence_after_sync();
            constexpr uint32_t idesc = make_idesc_tf32(128, kTK);
            constexpr uint32_t LBO = 128, SBO = (kTK / 4) * 128;
            const uint32_t a_hi = g.tcol + buf * (2 * kTKC), a_lo = a_hi + kTKC, d = g.tcol + dcol;
#pragma unroll
            for (int s = 0; s < kTKC / 8; ++s) {
                const int ks = c * (kTKC / 8) + s;
                const uint64_t bh = make_b_desc(b_hi + ks * 256, LBO, SBO);
                const uint64_t bl = make_b_desc(b_lo + ks * 256, LBO, SBO);
                mma_tf32_ts(d, a_lo + 8 * s, bh, idesc, ks > 0 ? 1u : 0u);
                mma_tf32_ts(d, a_hi + 8 * s, bl, idesc, 1u);
                mma_tf32_ts(d, a_hi + 8 * s, bh, idesc, 1u);
            }
            mma_commit(&g.pipe->empty[buf]);
            if (c == kTK / kTKC - 1) mma_commit(&g.pipe->dfull);
        }
    }
    __syncwarp();
    ++g.chunk;
}

__global__ void __launch_bounds__(2 * 128 + 32, 1)
tail_s1_kernel(const __grid_constant__ TailArgs p) {
    pdl_trigger();
    constexpr int K = kTK, KC = kTKC, NCH = K / KC;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) BPipe pipes[2];
    __shared__ __align__(8) uint64_t w1_full, w2_full, w1_free, w2_free;
    __shared__ __align__(16) float affs[2][2][2 * kTK];    // [set parity][pw1 | pw2][scale | shift]: BN terms outlive their pack's buffer
    __shared__ uint32_t tmem_slot;
    float* sB1 = smem;
    float* sB2 = sB1 + kTWFL;
    float* sDW = sB2 + kTWFL;
    float* Tb = sDW + K * 12;
    const int H = p.P.H, W = p.P.W, HW = H * W;
    const int TPL = (H + 2) * W + 2;                        // dense plane with a zero row above and below, 1 pad float each end
    const int warp = threadIdx.x >> 5;
    const int pairs = (p.N + 1) / 2;
    const int rounds = ((int)blockIdx.x < pairs) ? (pairs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nsets = rounds * p.nblk;                      // weight sets this CTA consumes

    if (threadIdx.x == 0) {
        mbar_init(&w1_full, 1); mbar_init(&w2_full, 1); mbar_init(&w1_free, 2); mbar_init(&w2_free, 2);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&pipes[i].empty[0], 1); mbar_init(&pipes[i].empty[1], 1); mbar_init(&pipes[i].dfull, 1);
            pipes[i].arrivals[0] = 0; pipes[i].arrivals[1] = 0;
        }
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    for (int i = threadIdx.x; i < 2 * 2 * KC * TPL; i += 2 * 128 + 32) Tb[i] = 0.f;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();

    if (warp == 8) {
        // ---------------- weight loader -------------------------------------------------------------------
        if ((threadIdx.x & 31) == 0) {
            for (int s = 0; s < nsets; ++s) {
                const int b = s % p.nblk;
                if (s > 0) { mbar_wait(&w1_free, (uint32_t)(s - 1) & 1u); publish_async(); }
                mbar_expect_tx(&w1_full, (uint32_t)(kTWFL * sizeof(float)));
                bulk_g2s(sB1, p.w1[b], kTWFL * sizeof(float), &w1_full);
                if (s > 0) { mbar_wait(&w2_free, (uint32_t)(s - 1) & 1u); publish_async(); }
                mbar_expect_tx(&w2_full, (uint32_t)((kTWFL + K * 12) * sizeof(float)));
                bulk_g2s(sB2, p.w2[b], kTWFL * sizeof(float), &w2_full);
                bulk_g2s(sDW, p.wdw[b], K * 12 * sizeof(float), &w2_full);
            }
        }
    } else {
        // ---------------- two warpgroups, one image each ----------------------------------------------------
        BGrp g;
        const int grp = threadIdx.x >> 7;
        g.tcol = tmem_slot + grp * kTCols;
        g.tlane = g.tcol + ((uint32_t)(32 * (warp & 3)) << 16);
        g.pipe = &pipes[grp];
        g.chunk = 0; g.dparity = 0;
        g.gtid = threadIdx.x & 127;
        const int q = g.gtid;
        const bool inpix = q < HW;
        const int qc = inpix ? q : 0;
        const int y = qc / W, x = qc - y * W;
        const float mL = x > 0 ? 1.f : 0.f, mR = x < W - 1 ? 1.f : 0.f;
        float* Tg = Tb + grp * (2 * KC * TPL);              // this group's two T chunk buffers
        const uint32_t b1_hi = smem_u32(sB1), b1_lo = smem_u32(sB1 + K * K);
        const uint32_t b2_hi = smem_u32(sB2), b2_lo = smem_u32(sB2 + K * K);
        pdl_wait();                                         // predecessor's activations are complete and visible from here on
        int s = 0;
        for (int r = 0; r < rounds; ++r) {
            const int n = 2 * ((int)blockIdx.x + r * (int)gridDim.x) + grp;
            const bool live = n < p.N;                      // an odd batch leaves the last round's second group without an image
            float* const pix = p.P.base + (long long)(live ? n : 0) * p.P.sN + p.P.org + y * p.P.Ws + x;
            for (int b = 0; b < p.nblk; ++b, ++s) {
                const uint32_t* ioff = p.in_off[b];
                const uint32_t* ooff = p.out_off[b];
                // ---- phase B: pw1 --------------------------------------------------------------------------
                float v[2][KC];
#pragma unroll
                for (int j = 0; j < KC; ++j) v[0][j] = __ldcg(pix + ioff[j]);
                mbar_wait(&w1_full, (uint32_t)s & 1u);
                float* aff1 = affs[s & 1][0];
                float* aff2 = affs[s & 1][1];
                for (int i = g.gtid; i < 2 * K; i += 128) aff1[i] = sB1[2 * K * K + i];
                group_bar(1 + grp, 128);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (c + 1 < NCH) {
#pragma unroll
                        for (int j = 0; j < KC; ++j) v[(c + 1) & 1][j] = __ldcg(pix + ioff[(c + 1) * KC + j]);
                    }
                    tail_acquire(g);
                    tail_store(g, v[c & 1]);
                    tail_hand_off(g, c, b1_hi, b1_lo, kTRing);
                }
                mbar_wait(&g.pipe->dfull, g.dparity);       // D1 complete: the pw1 pack is no longer read by this group
                g.dparity ^= 1u;
                fence_after_sync();
                mbar_wait(&w2_full, (uint32_t)s & 1u);
                for (int i = g.gtid; i < 2 * K; i += 128) aff2[i] = sB2[2 * K * K + i];      // read after the chunk loop's barriers
                // ---- phase C: D1 chunk -> T -> dw3x3 -> pw2 K-chunk ------------------------------------------
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    float* Tc = Tg + (c & 1) * (KC * TPL);
                    {
                        float d[KC];
                        tmem_ld16v(g.tlane + kTRing + c * KC, d);
                        wait_ld();
                        if (inpix) {
                            float* tp = Tc + 1 + (y + 1) * W + x;
#pragma unroll
                            for (int n4 = 0; n4 < KC; n4 += 4) {
                                const float4 sc = *reinterpret_cast<const float4*>(aff1 + c * KC + n4);
                                const float4 sh = *reinterpret_cast<const float4*>(aff1 + K + c * KC + n4);
                                tp[(n4 + 0) * TPL] = fmaxf(fmaf(d[n4 + 0], sc.x, sh.x), 0.f);
                                tp[(n4 + 1) * TPL] = fmaxf(fmaf(d[n4 + 1], sc.y, sh.y), 0.f);
                                tp[(n4 + 2) * TPL] = fmaxf(fmaf(d[n4 + 2], sc.z, sh.z), 0.f);
                                tp[(n4 + 3) * TPL] = fmaxf(fmaf(d[n4 + 3], sc.w, sh.w), 0.f);
                            }
                        }
                    }
                    if (c == 0 && (threadIdx.x & 127) == 0) mbar_arrive(&w1_free);     // (D1 is complete: see the wait above)
                    group_bar(1 + grp, 128);                // T chunk c complete; everyone is past the stencil of chunk c-2 (same buffer)
                    float a[KC];
                    {
                        const float* t = Tc + qc;           // window's top-left: row y (halo row 0 = image row -1), column x-1
                        const float* wk = sDW + c * KC * 12;
#pragma unroll
                        for (int jj = 0; jj < KC; ++jj) {
                            const float4 wa = *reinterpret_cast<const float4*>(wk);
                            const float4 wb = *reinterpret_cast<const float4*>(wk + 4);
                            const float4 wc = *reinterpret_cast<const float4*>(wk + 8);
                            float cl = wa.x * t[0]; cl = fmaf(wa.w, t[W], cl); cl = fmaf(wb.z, t[2 * W], cl);
                            float cc = wa.y * t[1]; cc = fmaf(wb.x, t[W + 1], cc); cc = fmaf(wb.w, t[2 * W + 1], cc);
                            float cr = wa.z * t[2]; cr = fmaf(wb.y, t[W + 2], cr); cr = fmaf(wc.x, t[2 * W + 2], cr);
                            const float dsum = fmaf(mR, cr, fmaf(mL, cl, cc));
                            a[jj] = fmaf(dsum, wc.y, wc.z);
                            t += TPL; wk += 12;
                        }
                    }
                    tail_acquire(g);
                    tail_store(g, a);
                    tail_hand_off(g, c, b2_hi, b2_lo, kTRing + K);
                }
                mbar_wait(&g.pipe->dfull, g.dparity);       // D2 complete: pw2 and dw packs are free
                g.dparity ^= 1u;
                fence_after_sync();
                if ((threadIdx.x & 127) == 0) mbar_arrive(&w2_free);
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    float d[KC];
                    tmem_ld16v(g.tlane + kTRing + K + c * KC, d);
                    wait_ld();
                    if (inpix && live) {
#pragma unroll
                        for (int n4 = 0; n4 < KC; n4 += 4) {
                            const float4 sc = *reinterpret_cast<const float4*>(aff2 + c * KC + n4);
                            const float4 sh = *reinterpret_cast<const float4*>(aff2 + K + c * KC + n4);
                            pix[ooff[c * KC + n4 + 0]] = fmaxf(fmaf(d[n4 + 0], sc.x, sh.x), 0.f);
                            pix[ooff[c * KC + n4 + 1]] = fmaxf(fmaf(d[n4 + 1], sc.y, sh.y), 0.f);
                            pix[ooff[c * KC + n4 + 2]] = fmaxf(fmaf(d[n4 + 2], sc.z, sh.z), 0.f);
                            pix[ooff[c * KC + n4 + 3]] = fmaxf(fmaf(d[n4 + 3], sc.w, sh.w), 0.f);
                        }
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

}  // namespace

// true if the K=96 stride-1 blocks of an H x W map can run on the small-map chain kernel
bool tail_s1_supported(int K, int H, int W) {
    if (K != kTK || H * W > 128) return false;
    const size_t bytes = ((size_t)2 * kTWFL + kTK * 12 + (size_t)2 * 2 * kTKC * ((H + 2) * W + 2)) * sizeof(float);
    return bytes <= kSmemCap - 1024;
}

int tail_launch_s1(const Planes& P, int nblk, const ChanTab* tin, const ChanTab* tout, const float* const* w1,
                   const float* const* wdw, const float* const* w2, int N, cudaStream_t s, int* done) {
    if (!tail_s1_supported(kTK, P.H, P.W)) { set_error("tail_launch_s1: unsupported map %dx%d", P.H, P.W); return YFV2_EUNSUPPORTED; }
    if (nblk > kTailMaxBlocks) nblk = kTailMaxBlocks;
    TailArgs a{};
    a.P = P; a.nblk = nblk; a.N = N;
    for (int b = 0; b < nblk; ++b) {
        a.w1[b] = w1[b]; a.w2[b] = w2[b]; a.wdw[b] = wdw[b];
        for (int k = 0; k < kTK; ++k) {
            a.in_off[b][k] = (uint32_t)((long long)tin[b].c[k] * P.sC);
            a.out_off[b][k] = (uint32_t)((long long)tout[b].c[k] * P.sC);
        }
    }
    const size_t bytes = ((size_t)2 * kTWFL + kTK * 12 + (size_t)2 * 2 * kTKC * ((P.H + 2) * P.W + 2)) * sizeof(float);
    YFV2_CUDA(cudaFuncSetAttribute(tail_s1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const int pairs = (N + 1) / 2;
    YFV2_CUDA(launch_k(tail_s1_kernel, min(pairs, sm_count()), 2 * 128 + 32, bytes, s, pdl_take(), a));
    YFV2_LAUNCH_CHECK();
    *done = nblk;
    return YFV2_OK;
}

}  // namespace yfv2
