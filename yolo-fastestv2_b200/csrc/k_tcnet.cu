// Tensor-core (tcgen05) versions of the fused network stages.  Every 1x1 convolution is a 3xTF32 UMMA
// (tc.cuh); depthwise 3x3 / 5x5, BN, ReLU, the stride-2 gather and all data movement stay on the CUDA cores.
//
// Execution model: one persistent CTA per SM made of G warpgroups.  A warpgroup (128 threads) owns one 128-row
// MMA tile at a time: thread t <-> pixel <-> TMEM lane.  It produces the A operand for its pixel in registers
// (global loads for a plain pointwise; the depthwise stencil for a DW->PW pair), splits it into tf32 hi/lo,
// stores it to TMEM, one elected thread issues the MMAs against the pre-tiled weights in shared memory, and
// after the commit barrier every thread pulls its output row back with tcgen05.ld for the epilogue.
// While one group waits for its MMAs the other groups run their CUDA-core phases.
//
// Kernels
//   tc_pw_kernel    out = act(BN(pw(in)))                       (FPN reducers with gather-on-load, K=96 pw1)
//   tc_dwpw_kernel  out = act(BN(pw( act(BN(dw(in))) )))        (+ optional chained output conv; heads, K=96 blocks)
//   tc_s1_kernel    ShuffleV2 stride-1 block, fully fused       (pw1 -> smem -> dw3x3 -> pw2)
//   tc_s2_kernel    ShuffleV2 stride-2 block, fully fused       (proj: dw s2 -> pw;  main: pw1 -> smem -> dw s2 -> pw2)
#include "common.cuh"
#include "tc.cuh"

namespace yfv2 {
namespace {

using namespace tc;

constexpr int kTmemCols = 512;

struct Grp {
    uint32_t tcol;     // TMEM address (lane 0) of this group's column block
    uint32_t tlane;    // same, at this warp's lane quarter
    uint64_t* mbar;
    uint32_t parity;
    int bar_id;
    int gtid;          // thread index inside the group, 0..127
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// One pointwise contraction for the group's current tile, K processed in KP/KC chunks (KC = KP: one shot).
//   load(k0, a[8])  : fill channels k0..k0+7 of this thread's pixel (zeros for padding channels / invalid pixels)
//   epi(n0, d[16])  : consume outputs n0..n0+15 of this thread's pixel
// TMEM columns of the group: A_hi [0,KC), A_lo [KC,2KC), D [2KC, 2KC+NP).
template <int KP, int NP, int KC, class Loader, class Epi>
__device__ __forceinline__ void pw_tile_c(Grp& g, uint32_t b_hi, uint32_t b_lo, Loader&& load, Epi&& epi) {
    static_assert(KP % KC == 0 && KC % 8 == 0, "chunking");
#pragma unroll 1
    for (int c = 0; c < KP / KC; ++c) {
        if (c > 0) {                         // the MMAs of the previous chunk must have consumed A before it is overwritten
            mbar_wait(g.mbar, g.parity);
            g.parity ^= 1u;
            fence_after_sync();
        }
#pragma unroll
        for (int k0 = 0; k0 < KC; k0 += 8) {
            float a[8];
            load(c * KC + k0, a);
            store_a8(g.tlane + k0, KC, a);
        }
        wait_st();
        fence_before_sync();
        group_bar(g.bar_id, 128);
        if (g.gtid == 0) {
            fence_after_sync();
            issue_pw<KC, NP, KP>(g.tcol + 2 * KC, g.tcol, g.tcol + KC, b_hi, b_lo, c * KC, c > 0);
            mma_commit(g.mbar);
        }
    }
    mbar_wait(g.mbar, g.parity);
    g.parity ^= 1u;
    fence_after_sync();
#pragma unroll
    for (int n0 = 0; n0 < NP; n0 += 16) {
        float d[16];
        tmem_ld16(g.tlane + 2 * KC + n0, d);
        wait_ld();
        epi(n0, d);
    }
    fence_before_sync();
    group_bar(g.bar_id, 128);      // D and A may be overwritten by the next tile from here on
}
template <int KP, int NP, class Loader, class Epi>
__device__ __forceinline__ void pw_tile(Grp& g, uint32_t b_hi, uint32_t b_lo, Loader&& load, Epi&& epi) {
    pw_tile_c<KP, NP, KP>(g, b_hi, b_lo, static_cast<Loader&&>(load), static_cast<Epi&&>(epi));
}

// CTA prologue shared by all kernels: TMEM allocation, barrier init, group context.
template <int G, int COLS>
__device__ __forceinline__ Grp cta_setup(uint64_t* mbars, uint32_t* tmem_slot) {
    static_assert(G * COLS <= kTmemCols, "TMEM columns");
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
    if (threadIdx.x == 32) {
        for (int i = 0; i < G; ++i) mbar_init(&mbars[i], 1);
        fence_mbar_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // weights written by this CTA -> visible to UMMA
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    Grp g;
    const int grp = warp >> 2;
    g.tcol = *tmem_slot + grp * COLS;
    g.tlane = g.tcol + ((uint32_t)(32 * (warp & 3)) << 16);
    g.mbar = &mbars[grp];
    g.parity = 0;
    g.bar_id = 1 + grp;
    g.gtid = threadIdx.x & 127;
    return g;
}
__device__ __forceinline__ void cta_teardown(uint32_t* tmem_slot) {
    fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc(*tmem_slot, kTmemCols);
}

__device__ __forceinline__ void copy_f4(float* dst, const float* __restrict__ src, int count, int nthreads) {
    for (int i = threadIdx.x * 4; i < count; i += nthreads * 4)
        *reinterpret_cast<float4*>(dst + i) = __ldg(reinterpret_cast<const float4*>(src + i));
}

// depthwise KSxKS (stride S) + BN (+ReLU) for 8 consecutive channels of one output pixel, from staged planes.
//   win: top-left of the window in plane 0 of the staged buffer; wdw: per-channel [KS*KS taps, scale, shift, pad]
template <int KS, int S, bool RELU>
__device__ __forceinline__ void dw8(const float* __restrict__ win, int RS, int WS, const float* __restrict__ wdw, int k0, bool valid,
                                    float (&a)[8]) {
    constexpr int R = KS == 3 ? 12 : 28;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float* wk = wdw + (k0 + j) * R;
        float w[R];
#pragma unroll
        for (int t = 0; t < R / 4; ++t) {
            const float4 w4 = *reinterpret_cast<const float4*>(wk + 4 * t);
            w[4 * t] = w4.x; w[4 * t + 1] = w4.y; w[4 * t + 2] = w4.z; w[4 * t + 3] = w4.w;
        }
        const float* xk = win + (k0 + j) * RS;
        float d = 0.f;
        if (valid) {
#pragma unroll
            for (int dy = 0; dy < KS; ++dy)
#pragma unroll
                for (int dx = 0; dx < KS; ++dx) d = fmaf(w[dy * KS + dx], xk[dy * WS + dx], d);
            d = fmaf(d, w[KS * KS], w[KS * KS + 1]);
            if (RELU) d = fmaxf(d, 0.f);
        }
        a[j] = d;
    }
}

// ===================================================================================================
// tc_pw_kernel: pointwise over planes.  KA channels read through (y>>SHA, x>>SHA) from A, then KB channels
// at (y,x) from B (K = KA+KB must be a multiple of 8).  One tile = 128 consecutive pixels of the flattened
// (image, pixel) index space.
// ===================================================================================================
struct PwArgs {
    Planes A, B, out;
    ChanTab ta, tb, tout;
    const float* wpack;     // tc pack: Bhi | Blo | scale | shift
    int N;                  // images
    int nout;               // real output channels
};

template <int KA, int KB, int SHA, int NP, int KC, int G, bool RELU>
__global__ void __launch_bounds__(G * 128, 1)
tc_pw_kernel(const __grid_constant__ PwArgs p) {
    constexpr int KP = KA + KB;
    static_assert(KP % 8 == 0 && NP % 16 == 0, "shape");
    constexpr int COLS = 2 * KC + NP;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t mbars[G];
    __shared__ uint32_t tmem_slot;
    float* sB = smem;
    constexpr int WFL = 2 * NP * KP + 2 * NP;
    copy_f4(sB, p.wpack, WFL, G * 128);
    Grp g = cta_setup<G, COLS>(mbars, &tmem_slot);
    const uint32_t b_hi = smem_u32(sB), b_lo = smem_u32(sB + NP * KP);
    const float* scale = sB + 2 * NP * KP;
    const float* shift = scale + NP;

    const int HW = p.out.H * p.out.W, W = p.out.W;
    const long long total = (long long)p.N * HW;
    const int ntiles = (int)((total + 127) / 128);
    const int grp = threadIdx.x >> 7;
    for (int tile = blockIdx.x * G + grp; tile < ntiles; tile += gridDim.x * G) {
        const long long pos = (long long)tile * 128 + g.gtid;
        const bool valid = pos < total;
        const int n = valid ? (int)(pos / HW) : 0;
        const int px = valid ? (int)(pos - (long long)n * HW) : 0;
        const int y = px / W, x = px - y * W;
        const float* baseA = p.A.base + (long long)n * p.A.sN + (long long)(y >> SHA) * p.A.W + (x >> SHA);
        const float* baseB = p.B.base + (long long)n * p.B.sN + px;
        float* obase = p.out.base + (long long)n * p.out.sN + px;
        pw_tile_c<KP, NP, KC>(g, b_hi, b_lo,
            [&](int k0, float (&a)[8]) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = k0 + j;
                    float v = 0.f;
                    if (valid) v = (k < KA) ? __ldg(baseA + (long long)p.ta.c[k] * p.A.sC) : __ldg(baseB + (long long)p.tb.c[k - KA] * p.B.sC);
                    a[j] = v;
                }
            },
            [&](int n0, float (&d)[16]) {
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int nn = n0 + j;
                        if (nn < p.nout) {
                            float v = fmaf(d[j], scale[nn], shift[nn]);
                            if (RELU) v = fmaxf(v, 0.f);
                            obase[(long long)p.tout.c[nn] * p.out.sC] = v;
                        }
                    }
                }
            });
    }
    cta_teardown(&tmem_slot);
}

// ===================================================================================================
// Band geometry shared by the DW-based kernels: a work item is (image, band of TR output rows).
// The staged buffer holds rows [S*r0 - PAD, S*(r0+rows-1) + PAD] of the source planes with PAD zero columns
// each side: staged row index rr <-> source row S*r0 - PAD + rr.
// ===================================================================================================
template <int KS, int S>
struct Band {
    static constexpr int PAD = KS / 2;
    __device__ static int staged_rows(int rows) { return S * (rows - 1) + KS; }
};

// DW(KSxKS, stride S)+BN(+ReLU) -> PW(KP->NP)+BN(+ReLU) -> store (optionally chained through a second PW).
struct DwPwArgs {
    Planes in[2], out[2];      // per branch
    ChanTab tin[2], tout[2];
    const float* wdw[2];       // DW pack per branch
    const float* wpw[2];       // tc pack per branch
    const float* wchain[2];    // chained output conv (tc pack, shift = bias), CHAIN only
    float* dstA[2]; float* dstB[2]; int split[2]; int M[2];   // CHAIN: dense NCHW destinations
    int N, TR, bandsPerImg, nbranch, nout;
};

template <int K, int NP, int G, int KS, int S, bool RELU_DW, bool RELU_OUT, bool CHAIN, int NP2>
__global__ void __launch_bounds__(G * 128, 1)
tc_dwpw_kernel(const __grid_constant__ DwPwArgs p) {
    constexpr int KP = K;
    static_assert(KP % 8 == 0 && NP % 16 == 0, "shape");
    constexpr int COLS = CHAIN ? (2 * KP + NP + ((2 * NP + NP2 > 2 * KP + NP) ? (2 * NP + NP2 - 2 * KP - NP) : 0)) : 2 * KP + NP;
    constexpr int PAD = KS / 2;
    constexpr int DWR = KS == 3 ? 12 : 28;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t mbars[G];
    __shared__ uint32_t tmem_slot;
    constexpr int WFL = 2 * NP * KP + 2 * NP;
    constexpr int WFL2 = CHAIN ? 2 * NP2 * NP + 2 * NP2 : 0;
    float* sB = smem;                          // pw pack
    float* sB2 = sB + WFL;                     // chained pack
    float* sDW = sB2 + WFL2;                   // dw pack
    float* X = sDW + K * DWR;                  // staged planes

    const int Hout = p.out[0].H, Wout = p.out[0].W;
    const int Win = p.in[0].W;
    const int WS = Win + 2 * PAD;
    const int RS = (S * (p.TR - 1) + KS) * WS;
    Grp g = cta_setup<G, COLS>(mbars, &tmem_slot);
    const uint32_t b_hi = smem_u32(sB), b_lo = smem_u32(sB + NP * KP);
    const uint32_t c_hi = smem_u32(sB2), c_lo = smem_u32(sB2 + NP2 * NP);
    const float* scale = sB + 2 * NP * KP;
    const float* shift = scale + NP;
    const float* bias2 = sB2 + 2 * NP2 * NP + NP2;
    const int grp = threadIdx.x >> 7;
    const int items = p.N * p.bandsPerImg * p.nbranch;
    int loaded_branch = -1;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int br = item % p.nbranch;
        const int rem = item / p.nbranch;
        const int n = rem / p.bandsPerImg;
        const int r0 = (rem - n * p.bandsPerImg) * p.TR;
        const int rows = min(p.TR, Hout - r0);
        __syncthreads();                                   // previous item done with X / weights
        if (br != loaded_branch) {
            copy_f4(sB, p.wpw[br], WFL, G * 128);
            if (CHAIN) copy_f4(sB2, p.wchain[br], WFL2, G * 128);
            copy_f4(sDW, p.wdw[br], K * DWR, G * 128);
            loaded_branch = br;
        }
        stage_rows<K, PAD, G * 128>(X, RS, WS, p.in[br], p.tin[br], n, S * r0 - PAD, S * (rows - 1) + KS);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        const int npix = rows * Wout;
        const int ntiles = (npix + 127) / 128;
        for (int tile = grp; tile < ntiles; tile += G) {
            const int q = tile * 128 + g.gtid;
            const bool valid = q < npix;
            const int orow = valid ? q / Wout : 0, ox = valid ? q - orow * Wout : 0;
            const float* win = X + (S * orow) * WS + S * ox;
            const long long opix = (long long)(r0 + orow) * Wout + ox;
            if (!CHAIN) {
                float* obase = p.out[br].base + (long long)n * p.out[br].sN + opix;
                pw_tile<KP, NP>(g, b_hi, b_lo,
                    [&](int k0, float (&a)[8]) { dw8<KS, S, RELU_DW>(win, RS, WS, sDW, k0, valid, a); },
                    [&](int n0, float (&d)[16]) {
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int nn = n0 + j;
                                if (nn < p.nout) {
                                    float v = fmaf(d[j], scale[nn], shift[nn]);
                                    if (RELU_OUT) v = fmaxf(v, 0.f);
                                    obase[(long long)p.tout[br].c[nn] * p.out[br].sC] = v;
                                }
                            }
                        }
                    });
            } else {
                // features (NP columns, real p.nout) -> BN -> second UMMA against the output conv -> dense NCHW
                // TMEM of the second contraction: A_hi [0,NP) A_lo [NP,2NP) D2 [2NP, 2NP+NP2) — it reuses the
                // group's columns once the first D has been read back.
                float f[NP];
                pw_tile<KP, NP>(g, b_hi, b_lo,
                    [&](int k0, float (&a)[8]) { dw8<KS, S, RELU_DW>(win, RS, WS, sDW, k0, valid, a); },
                    [&](int n0, float (&d)[16]) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float v = fmaf(d[j], scale[n0 + j], shift[n0 + j]);
                            if (RELU_OUT) v = fmaxf(v, 0.f);
                            f[n0 + j] = valid ? v : 0.f;
                        }
                    });
                const int HW = Hout * Wout;
                const int split = p.split[br], M = p.M[br];
                float* dA = p.dstA[br]; float* dB = p.dstB[br];
                pw_tile<NP, NP2>(g, c_hi, c_lo,
                    [&](int k0, float (&a)[8]) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) a[j] = f[k0 + j];
                    },
                    [&](int n0, float (&d)[16]) {
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int m = n0 + j;
                                const float v = d[j] + bias2[m];
                                if (m < split) dA[((long long)n * split + m) * HW + opix] = v;
                                else if (m < M) dB[((long long)n * (M - split) + (m - split)) * HW + opix] = v;
                            }
                        }
                    });
            }
        }
    }
    cta_teardown(&tmem_slot);
}

// ===================================================================================================
// tc_s1_kernel: fused stride-1 ShuffleV2 block (reference shufflenetv2.py:19-32,48-51).
// ===================================================================================================
struct S1Args {
    Planes P;
    ChanTab tin, tout;
    const float* w1;     // tc pack pw1
    const float* wdw;    // dw3 pack
    const float* w2;     // tc pack pw2
    int N, TR, bandsPerImg;
};

template <int K, int NP, int G>
__global__ void __launch_bounds__(G * 128, 1)
tc_s1_kernel(const __grid_constant__ S1Args p) {
    constexpr int KP = K;
    constexpr int COLS = 2 * KP + NP;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t mbars[G];
    __shared__ uint32_t tmem_slot;
    constexpr int WFL = 2 * NP * KP + 2 * NP;
    float* sB1 = smem;
    float* sB2 = sB1 + WFL;
    float* sDW = sB2 + WFL;
    float* T = sDW + K * 12;
    const int H = p.P.H, W = p.P.W, WS = W + 2;
    const int RS = (p.TR + 2) * WS;
    copy_f4(sB1, p.w1, WFL, G * 128);
    copy_f4(sB2, p.w2, WFL, G * 128);
    copy_f4(sDW, p.wdw, K * 12, G * 128);
    Grp g = cta_setup<G, COLS>(mbars, &tmem_slot);
    const uint32_t b1_hi = smem_u32(sB1), b1_lo = smem_u32(sB1 + NP * KP);
    const uint32_t b2_hi = smem_u32(sB2), b2_lo = smem_u32(sB2 + NP * KP);
    const float* sc1 = sB1 + 2 * NP * KP; const float* sh1 = sc1 + NP;
    const float* sc2 = sB2 + 2 * NP * KP; const float* sh2 = sc2 + NP;
    const int grp = threadIdx.x >> 7;
    const int items = p.N * p.bandsPerImg;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int n = item / p.bandsPerImg;
        const int r0 = (item - n * p.bandsPerImg) * p.TR;
        const int rows = min(p.TR, H - r0);
        __syncthreads();
        // zero the band buffer: padding columns / out-of-image rows must read as 0 for the depthwise
        for (int i = threadIdx.x * 4; i < K * RS; i += G * 128 * 4) *reinterpret_cast<float4*>(T + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        // ---- phase B: pw1 + BN + ReLU on every in-image pixel of rows [r0-1, r0+rows] -> T --------------------------
        const int gr_lo = max(r0 - 1, 0), gr_hi = min(r0 + rows, H - 1);
        const int npos = (gr_hi - gr_lo + 1) * W;
        for (int tile = grp; tile * 128 < npos; tile += G) {
            const int q = tile * 128 + g.gtid;
            const bool valid = q < npos;
            const int rr = valid ? q / W : 0, x = valid ? q - rr * W : 0;
            const int gr = gr_lo + rr;
            const float* ibase = p.P.base + (long long)n * p.P.sN + (long long)gr * W + x;
            float* tpos = T + (gr - (r0 - 1)) * WS + 1 + x;
            pw_tile<KP, NP>(g, b1_hi, b1_lo,
                [&](int k0, float (&a)[8]) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = valid ? __ldg(ibase + (long long)p.tin.c[k0 + j] * p.P.sC) : 0.f;
                },
                [&](int n0, float (&d)[16]) {
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + j < K) tpos[(n0 + j) * RS] = fmaxf(fmaf(d[j], sc1[n0 + j], sh1[n0 + j]), 0.f);
                    }
                });
        }
        __syncthreads();
        // ---- phase C: dw3x3 + BN -> pw2 + BN + ReLU -> output planes -------------------------------------------
        const int npix = rows * W;
        for (int tile = grp; tile * 128 < npix; tile += G) {
            const int q = tile * 128 + g.gtid;
            const bool valid = q < npix;
            const int orow = valid ? q / W : 0, ox = valid ? q - orow * W : 0;
            const float* win = T + orow * WS + ox;
            float* obase = p.P.base + (long long)n * p.P.sN + (long long)(r0 + orow) * W + ox;
            pw_tile<KP, NP>(g, b2_hi, b2_lo,
                [&](int k0, float (&a)[8]) { dw8<3, 1, false>(win, RS, WS, sDW, k0, valid, a); },
                [&](int n0, float (&d)[16]) {
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + j < K) obase[(long long)p.tout.c[n0 + j] * p.P.sC] = fmaxf(fmaf(d[j], sc2[n0 + j], sh2[n0 + j]), 0.f);
                    }
                });
        }
    }
    cta_teardown(&tmem_slot);
}

// ===================================================================================================
// tc_s2_kernel: fused stride-2 ShuffleV2 block (reference shufflenetv2.py:34-44,52-55).
// ===================================================================================================
struct S2Args {
    Planes in, out;
    ChanTab tin, tout;
    const float* wdwp; const float* wp;      // proj: dw pack, tc pack
    const float* w1; const float* wdwm; const float* w2;
    int N, TR, bandsPerImg;
};

template <int K, int NP, int G>
__global__ void __launch_bounds__(G * 128, 1)
tc_s2_kernel(const __grid_constant__ S2Args p) {
    constexpr int KP = K;
    constexpr int COLS = 2 * KP + NP;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t mbars[G];
    __shared__ uint32_t tmem_slot;
    constexpr int WFL = 2 * NP * KP + 2 * NP;
    float* sBp = smem;
    float* sB1 = sBp + WFL;
    float* sB2 = sB1 + WFL;
    float* sDWp = sB2 + WFL;
    float* sDWm = sDWp + K * 12;
    float* X = sDWm + K * 12;
    const int Hin = p.in.H, Win = p.in.W, WS = Win + 2;
    const int Hout = p.out.H, Wout = p.out.W;
    const int RS = (2 * p.TR + 1) * WS;
    copy_f4(sBp, p.wp, WFL, G * 128);
    copy_f4(sB1, p.w1, WFL, G * 128);
    copy_f4(sB2, p.w2, WFL, G * 128);
    copy_f4(sDWp, p.wdwp, K * 12, G * 128);
    copy_f4(sDWm, p.wdwm, K * 12, G * 128);
    Grp g = cta_setup<G, COLS>(mbars, &tmem_slot);
    const uint32_t bp_hi = smem_u32(sBp), bp_lo = smem_u32(sBp + NP * KP);
    const uint32_t b1_hi = smem_u32(sB1), b1_lo = smem_u32(sB1 + NP * KP);
    const uint32_t b2_hi = smem_u32(sB2), b2_lo = smem_u32(sB2 + NP * KP);
    const float* scp = sBp + 2 * NP * KP; const float* shp = scp + NP;
    const float* sc1 = sB1 + 2 * NP * KP; const float* sh1 = sc1 + NP;
    const float* sc2 = sB2 + 2 * NP * KP; const float* sh2 = sc2 + NP;
    const int grp = threadIdx.x >> 7;
    const int items = p.N * p.bandsPerImg;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int n = item / p.bandsPerImg;
        const int r0 = (item - n * p.bandsPerImg) * p.TR;
        const int rows = min(p.TR, Hout - r0);
        const int gr0 = 2 * r0 - 1, nrows = 2 * rows + 1;
        __syncthreads();
        stage_rows<K, 1, G * 128>(X, RS, WS, p.in, p.tin, n, gr0, nrows);
        __syncthreads();
        const int npix = rows * Wout;
        // ---- proj: dw3x3 s2 + BN -> pw + BN + ReLU on the raw input ---------------------------------------------
        for (int tile = grp; tile * 128 < npix; tile += G) {
            const int q = tile * 128 + g.gtid;
            const bool valid = q < npix;
            const int orow = valid ? q / Wout : 0, ox = valid ? q - orow * Wout : 0;
            const float* win = X + (2 * orow) * WS + 2 * ox;
            float* obase = p.out.base + (long long)n * p.out.sN + (long long)(r0 + orow) * Wout + ox;
            pw_tile<KP, NP>(g, bp_hi, bp_lo,
                [&](int k0, float (&a)[8]) { dw8<3, 2, false>(win, RS, WS, sDWp, k0, valid, a); },
                [&](int n0, float (&d)[16]) {
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + j < K) obase[(long long)p.tout.c[n0 + j] * p.out.sC] = fmaxf(fmaf(d[j], scp[n0 + j], shp[n0 + j]), 0.f);
                    }
                });
        }
        __syncthreads();
        // ---- main pw1 in place on every staged in-image pixel ------------------------------------------------------
        const int gr_lo = max(gr0, 0), gr_hi = min(gr0 + nrows - 1, Hin - 1);
        const int npos = (gr_hi - gr_lo + 1) * Win;
        for (int tile = grp; tile * 128 < npos; tile += G) {
            const int q = tile * 128 + g.gtid;
            const bool valid = q < npos;
            const int rr = valid ? q / Win : 0, x = valid ? q - rr * Win : 0;
            float* tpos = X + (gr_lo + rr - gr0) * WS + 1 + x;
            pw_tile<KP, NP>(g, b1_hi, b1_lo,
                [&](int k0, float (&a)[8]) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = valid ? tpos[(k0 + j) * RS] : 0.f;
                },
                [&](int n0, float (&d)[16]) {
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + j < K) tpos[(n0 + j) * RS] = fmaxf(fmaf(d[j], sc1[n0 + j], sh1[n0 + j]), 0.f);
                    }
                });
        }
        __syncthreads();
        // ---- main: dw3x3 s2 + BN -> pw2 + BN + ReLU ---------------------------------------------------------------
        for (int tile = grp; tile * 128 < npix; tile += G) {
            const int q = tile * 128 + g.gtid;
            const bool valid = q < npix;
            const int orow = valid ? q / Wout : 0, ox = valid ? q - orow * Wout : 0;
            const float* win = X + (2 * orow) * WS + 2 * ox;
            float* obase = p.out.base + (long long)n * p.out.sN + (long long)(r0 + orow) * Wout + ox;
            pw_tile<KP, NP>(g, b2_hi, b2_lo,
                [&](int k0, float (&a)[8]) { dw8<3, 2, false>(win, RS, WS, sDWm, k0, valid, a); },
                [&](int n0, float (&d)[16]) {
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + j < K) obase[(long long)p.tout.c[K + n0 + j] * p.out.sC] = fmaxf(fmaf(d[j], sc2[n0 + j], sh2[n0 + j]), 0.f);
                    }
                });
        }
    }
    cta_teardown(&tmem_slot);
}

template <typename Kern>
int set_smem_attr(Kern kern, size_t bytes) {
    if (bytes > kSmemCap) { set_error("tc kernel needs %zu bytes of shared memory", bytes); return YFV2_EUNSUPPORTED; }
    YFV2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return YFV2_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
#define TRYL(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

int tc_launch_s1(int K, const Planes& P, const ChanTab& tin, const ChanTab& tout, const float* w1, const float* wdw, const float* w2,
                 int N, cudaStream_t s) {
    S1Args a{P, tin, tout, w1, wdw, w2, N, 0, 0};
    const int H = P.H, W = P.W;
    auto run = [&](auto kern, int KK, int NP, int G) -> int {
        const size_t wfl = (size_t)2 * (2 * NP * KK + 2 * NP) + KK * 12;
        auto bytes = [&](int tr) { return (wfl + (size_t)KK * (tr + 2) * (W + 2) + 4) * sizeof(float); };
        int TR = H;
        while (TR > 1 && bytes(TR) > 200 * 1024) TR = (TR + 1) / 2;
        a.TR = TR; a.bandsPerImg = (H + TR - 1) / TR;
        TRYL(set_smem_attr(kern, bytes(TR)));
        const int items = N * a.bandsPerImg;
        kern<<<min(items, sm_count()), G * 128, bytes(TR), s>>>(a);
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    if (K == 24) return run(tc_s1_kernel<24, 32, 4>, 24, 32, 4);
    if (K == 48) return run(tc_s1_kernel<48, 48, 3>, 48, 48, 3);
    set_error("tc_launch_s1: unsupported K=%d", K);
    return YFV2_EUNSUPPORTED;
}

int tc_launch_s2(int K, const Planes& in, const Planes& out, const ChanTab& tin, const ChanTab& tout, const float* wdwp, const float* wp,
                 const float* w1, const float* wdwm, const float* w2, int N, cudaStream_t s) {
    S2Args a{in, out, tin, tout, wdwp, wp, w1, wdwm, w2, N, 0, 0};
    const int Hout = out.H, Win = in.W;
    auto run = [&](auto kern, int KK, int NP, int G) -> int {
        const size_t wfl = (size_t)3 * (2 * NP * KK + 2 * NP) + 2 * KK * 12;
        auto bytes = [&](int tr) { return (wfl + (size_t)KK * (2 * tr + 1) * (Win + 2) + 4) * sizeof(float); };
        int TR = Hout;
        while (TR > 1 && bytes(TR) > 200 * 1024) --TR;
        a.TR = TR; a.bandsPerImg = (Hout + TR - 1) / TR;
        TRYL(set_smem_attr(kern, bytes(TR)));
        const int items = N * a.bandsPerImg;
        kern<<<min(items, sm_count()), G * 128, bytes(TR), s>>>(a);
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    if (K == 24) return run(tc_s2_kernel<24, 32, 4>, 24, 32, 4);
    if (K == 48) return run(tc_s2_kernel<48, 48, 3>, 48, 48, 3);
    set_error("tc_launch_s2: unsupported K=%d", K);
    return YFV2_EUNSUPPORTED;
}

// plain pointwise; kind 0: 96->96 (+ReLU)  1: FPN S3 192->72 (+ReLU)  2: FPN S2 (up(192) ++ 96)->72 (+ReLU)
int tc_launch_pw(int kind, const Planes& A, const ChanTab& ta, const Planes& B, const ChanTab& tb, const Planes& out, const ChanTab& tout,
                 const float* wpack, int N, cudaStream_t s) {
    PwArgs a{A, B, out, ta, tb, tout, wpack, N, 0};
    const long long total = (long long)N * out.H * out.W;
    const int ntiles = (int)((total + 127) / 128);
    auto run = [&](auto kern, int KP, int NP, int G, int nout) -> int {
        a.nout = nout;
        const size_t bytes = (size_t)(2 * NP * KP + 2 * NP) * sizeof(float);
        TRYL(set_smem_attr(kern, bytes));
        kern<<<min((ntiles + G - 1) / G, sm_count()), G * 128, bytes, s>>>(a);
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    if (kind == 0) return run(tc_pw_kernel<96, 0, 0, 96, 96, 1, true>, 96, 96, 1, 96);
    if (kind == 1) return run(tc_pw_kernel<192, 0, 0, 80, 96, 1, true>, 192, 80, 1, 72);
    if (kind == 2) return run(tc_pw_kernel<192, 96, 1, 80, 96, 1, true>, 288, 80, 1, 72);
    set_error("tc_launch_pw: unknown kind %d", kind);
    return YFV2_EUNSUPPORTED;
}

// K=96 DW3x3(stride)->PW (+ReLU) for nbranch branches (stage-4 blocks)
int tc_launch_dwpw96(int stride, int nbranch, const Planes* in, const ChanTab* tin, const Planes* out, const ChanTab* tout,
                     const float* const* wdw, const float* const* wpw, int N, cudaStream_t s) {
    DwPwArgs a{};
    for (int b = 0; b < nbranch; ++b) { a.in[b] = in[b]; a.out[b] = out[b]; a.tin[b] = tin[b]; a.tout[b] = tout[b]; a.wdw[b] = wdw[b]; a.wpw[b] = wpw[b]; }
    a.N = N; a.nbranch = nbranch; a.nout = 96;
    const int Hout = out[0].H, Win = in[0].W;
    auto run = [&](auto kern, int S) -> int {
        const size_t wfl = (size_t)(2 * 96 * 96 + 2 * 96) + 96 * 12;
        auto bytes = [&](int tr) { return (wfl + (size_t)96 * (S * (tr - 1) + 3) * (Win + 2) + 4) * sizeof(float); };
        int TR = Hout;
        while (TR > 1 && bytes(TR) > 200 * 1024) --TR;
        a.TR = TR; a.bandsPerImg = (Hout + TR - 1) / TR;
        TRYL(set_smem_attr(kern, bytes(TR)));
        const int items = N * a.bandsPerImg * nbranch;
        kern<<<min(items, sm_count()), 128, bytes(TR), s>>>(a);
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    if (stride == 1) return run(tc_dwpw_kernel<96, 96, 1, 3, 1, false, true, false, 16>, 1);
    return run(tc_dwpw_kernel<96, 96, 1, 3, 2, false, true, false, 16>, 2);
}

// heads: half 0: T = BN(pw(ReLU(BN(dw5x5(S)))));  half 1: preds = outconv(BN(pw(ReLU(BN(dw5x5(T)))))) + bias.
// branch 0 = cls head (outputs obj+cls), branch 1 = reg head.
int tc_launch_heads(int half, const Planes& sIn, const Planes& tcls, const Planes& treg, const float* const wdw[2], const float* const wpw[2],
                    const float* wout_oc, const float* wout_reg, float* reg, float* obj, float* cls, int A, int C, int N, cudaStream_t s) {
    DwPwArgs a{};
    ChanTab ident;
    for (int i = 0; i < kMaxCh; ++i) ident.c[i] = (unsigned short)i;
    a.N = N; a.nbranch = 2; a.nout = 72;
    for (int b = 0; b < 2; ++b) { a.tin[b] = ident; a.tout[b] = ident; a.wdw[b] = wdw[b]; a.wpw[b] = wpw[b]; }
    if (half == 0) { a.in[0] = sIn; a.in[1] = sIn; a.out[0] = tcls; a.out[1] = treg; }
    else {
        a.in[0] = tcls; a.in[1] = treg; a.out[0] = tcls; a.out[1] = treg;
        a.wchain[0] = wout_oc; a.wchain[1] = wout_reg;
        a.dstA[0] = obj; a.dstB[0] = cls; a.split[0] = A; a.M[0] = A + C;
        a.dstA[1] = reg; a.dstB[1] = reg; a.split[1] = 4 * A; a.M[1] = 4 * A;
    }
    const int Hout = sIn.H, Win = sIn.W;
    constexpr int NP = 80, NP2 = 96;
    if (A + C > NP2 || 4 * A > NP2) { set_error("tc heads: A+C=%d exceeds the chained tile (%d)", A + C, NP2); return YFV2_EUNSUPPORTED; }
    auto run = [&](auto kern, bool chain) -> int {
        const size_t wfl = (size_t)(2 * NP * 72 + 2 * NP) + (chain ? (size_t)(2 * NP2 * NP + 2 * NP2) : 0) + 72 * 28;
        auto bytes = [&](int tr) { return (wfl + (size_t)72 * (tr + 4) * (Win + 4) + 4) * sizeof(float); };
        int TR = Hout;
        while (TR > 1 && bytes(TR) > 200 * 1024) --TR;
        a.TR = TR; a.bandsPerImg = (Hout + TR - 1) / TR;
        TRYL(set_smem_attr(kern, bytes(TR)));
        const int items = N * a.bandsPerImg * 2;
        kern<<<min(items, sm_count()), 2 * 128, bytes(TR), s>>>(a);
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    if (half == 0) return run(tc_dwpw_kernel<72, NP, 2, 5, 1, true, false, false, 16>, false);
    return run(tc_dwpw_kernel<72, NP, 2, 5, 1, true, false, true, NP2>, true);
}

}  // namespace yfv2
