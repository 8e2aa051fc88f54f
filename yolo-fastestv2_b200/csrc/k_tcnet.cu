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
//
// Staging: planes live in global memory inside zero frames (common.cuh), so the rows a band needs, halo and
// padding included, are one contiguous 16-byte-aligned run per plane: one TMA bulk copy (cp.async.bulk, UBLKCP)
// per plane, issued by K different threads, completion counted on an mbarrier.  No per-element address math.
#include <string.h>

#include <vector>

#include "common.cuh"
#include "tc.cuh"

namespace yfv2 {
namespace {

using namespace tc;

// ---------------------------------------------------------------------------------------------------------
// Engine.  A CTA = G warpgroups (128 threads, thread <-> pixel <-> TMEM lane).  The A operand streams through TMEM
// in 8-channel chunks, double buffered, and there is NO dedicated issuer: whichever of the group's four warps is
// the last to finish storing a chunk issues that chunk's MMAs (an acq_rel shared-memory counter decides), so no
// warp ever waits for an MMA to be *issued*:
//   every warp:  (wait empty[buf]) -> split 8 channels -> tcgen05.st hi/lo -> wait::st, fence -> counter++
//   4th arriver: 3 x tcgen05.mma (K=8) -> tcgen05.commit -> empty[buf]   (+ commit -> dfull after the tile's last chunk)
//   every warp:  wait dfull -> tcgen05.ld D -> epilogue
// A group needs only 32 + NP TMEM columns whatever K is; the MMAs of a chunk overlap the depthwise / load work of
// the following chunks and of the other groups.  Two such CTAs fit on an SM (each allocates 256 of the 512 columns).
// ---------------------------------------------------------------------------------------------------------
constexpr int kTmemCols = 256;          // per CTA; every kernel here allocates the same amount (two CTAs fill an SM's 512)
constexpr int kACols = 32;             // 2 buffers x (8 hi + 8 lo)

constexpr int kMaxABufs = 4;
struct Pipe {                          // per group, in shared memory
    uint64_t empty[kMaxABufs], dfull;
    uint32_t arrivals[kMaxABufs];
    uint32_t pad_[2];
};

struct Grp {
    uint32_t tcol;                     // TMEM address (lane 0) of the group's column block
    uint32_t tlane;                    // same, at this warp's lane quarter
    Pipe* pipe;
    uint32_t chunk;                    // chunks produced so far (buffer = chunk & 1, use index = chunk >> 1)
    uint32_t dparity;
    int gtid;                          // thread index inside the group, 0..127
};

__device__ __forceinline__ uint32_t atom_add_acq_rel(uint32_t* addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(addr)), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Hand chunk c (8 channels of this thread's pixel) of a KP->NP contraction to the tensor core.
//   b_hi / b_lo: shared addresses of the weight pack; `last`: this is the tile's final chunk.
//   NB: A buffers in the ring (2 everywhere; the accumulator then starts at column 16*NB of the group's block)
template <int KP, int NP, int NB = 2>
__device__ __forceinline__ void put_chunk(Grp& g, const float (&a)[8], int c, uint32_t b_hi, uint32_t b_lo) {
    static_assert(NB >= 2 && NB <= kMaxABufs, "A ring depth");
    const uint32_t buf = g.chunk % NB, use = g.chunk / NB;
    if (use > 0) mbar_wait(&g.pipe->empty[buf], (use - 1) & 1u);     // the MMAs that read this buffer last time are done
    fence_after_sync();
    store_a8(g.tlane + buf * 16, 8, a);
    wait_st();
    fence_before_sync();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const uint32_t old = atom_add_acq_rel(&g.pipe->arrivals[buf], 1u);
        if ((old & 3u) == 3u) {                                      // all four warps of the group have stored this chunk
            fence_after_sync();
            constexpr uint32_t idesc = make_idesc_tf32(128, NP);
            constexpr uint32_t LBO = 128, SBO = (KP / 4) * 128;
            const uint32_t a_hi = g.tcol + buf * 16, a_lo = a_hi + 8, d = g.tcol + 16 * NB;
            const uint64_t bh = make_b_desc(b_hi + c * 256, LBO, SBO);
            const uint64_t bl = make_b_desc(b_lo + c * 256, LBO, SBO);
            mma_tf32_ts(d, a_lo, bh, idesc, c > 0 ? 1u : 0u);        // small terms first
            mma_tf32_ts(d, a_hi, bl, idesc, 1u);
            mma_tf32_ts(d, a_hi, bh, idesc, 1u);
            mma_commit(&g.pipe->empty[buf]);
            if (c == KP / 8 - 1) mma_commit(&g.pipe->dfull);
        }
    }
    __syncwarp();
    ++g.chunk;
}
// Collect the NP output columns of this thread's pixel.
template <int NP, int NB = 2, class Epi>
__device__ __forceinline__ void get_tile(Grp& g, Epi&& epi, int ncols = NP) {
    mbar_wait(&g.pipe->dfull, g.dparity);
    g.dparity ^= 1u;
    fence_after_sync();
#pragma unroll
    for (int n0 = 0; n0 < NP; n0 += 16) {
        if (n0 >= ncols) break;                 // trailing all-padding column blocks (warp-uniform)
        float d[16];
        tmem_ld16(g.tlane + 16 * NB + n0, d);
        wait_ld();
        epi(n0, d);
    }
    // the next put_chunk's fence::before_thread_sync + counter increment orders these reads before D is overwritten
}
template <int KP, int NP, class Loader, class Epi>
__device__ __forceinline__ void pw_tile(Grp& g, uint32_t b_hi, uint32_t b_lo, Loader&& load, Epi&& epi) {
#pragma unroll
    for (int k0 = 0; k0 < KP; k0 += 8) {
        float a[8];
        load(k0, a);
        put_chunk<KP, NP>(g, a, k0 / 8, b_hi, b_lo);
    }
    get_tile<NP>(g, static_cast<Epi&&>(epi));
}

// Same with the chunk loop NOT unrolled: for loaders that accept a runtime k0 (the depthwise stencils).  The fully
// unrolled 72-channel 5x5 stencil is ~100 KB of SASS and starves the instruction cache with 8 warps per SM.
template <int KP, int NP, class Loader, class Epi>
__device__ __forceinline__ void pw_tile_rolled(Grp& g, uint32_t b_hi, uint32_t b_lo, Loader&& load, Epi&& epi) {
#pragma unroll 1
    for (int k0 = 0; k0 < KP; k0 += 8) {
        float a[8];
        load(k0, a);
        put_chunk<KP, NP>(g, a, k0 >> 3, b_hi, b_lo);
    }
    get_tile<NP>(g, static_cast<Epi&&>(epi));
}

// CTA prologue: TMEM allocation, barrier init.
template <int G, int COLS, int TOT = kTmemCols>
__device__ __forceinline__ Grp cta_setup(Pipe* pipes, uint32_t* tmem_slot) {
    static_assert(G * COLS <= TOT, "TMEM columns");
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(tmem_slot, TOT);
    if (threadIdx.x == 32) {
        for (int i = 0; i < G; ++i) {
            for (int b = 0; b < kMaxABufs; ++b) { mbar_init(&pipes[i].empty[b], 1); pipes[i].arrivals[b] = 0; }
            mbar_init(&pipes[i].dfull, 1);
        }
        fence_mbar_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const int grp = threadIdx.x >> 7;
    Grp g;
    g.tcol = *tmem_slot + grp * COLS;
    g.tlane = g.tcol + ((uint32_t)(32 * (warp & 3)) << 16);
    g.pipe = &pipes[grp];
    g.chunk = 0; g.dparity = 0;
    g.gtid = threadIdx.x & 127;
    return g;
}
template <int TOT = kTmemCols>
__device__ __forceinline__ void cta_teardown(uint32_t* tmem_slot) {
    fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc(*tmem_slot, TOT);
}
template <int G>
__device__ __forceinline__ void producers_sync() { __syncthreads(); }
// make weights written with ordinary stores visible to the tensor core's async proxy
__device__ __forceinline__ void publish_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Stage `nrows` frame rows starting at frame row fr0 of the K planes in `tab` (image n) into X[k][.] (plane stride RS
// floats).  One bulk copy per plane.  Callers: all threads; wait with mbar_wait(bar, parity) afterwards.
template <int K>
__device__ __forceinline__ void stage_bulk(float* X, int RS, const Planes& P, const ChanTab& tab, int n, int fr0, int nrows, uint64_t* bar,
                                           int tid0) {
    const int t = (int)threadIdx.x - tid0;
    if (t >= 0 && t < K) {
        const float* src = plane_ptr(P, n, tab.c[t]) + (long long)fr0 * P.Ws;
        bulk_g2s(X + (size_t)t * RS, src, (uint32_t)(nrows * P.Ws * sizeof(float)), bar);
    }
}

__device__ __forceinline__ void copy_f4(float* dst, const float* __restrict__ src, int count, int nthreads) {
    for (int i = threadIdx.x * 4; i < count; i += nthreads * 4)
        *reinterpret_cast<float4*>(dst + i) = __ldg(reinterpret_cast<const float4*>(src + i));
}

// depthwise KSxKS (stride S) + BN (+ReLU) for 8 consecutive channels of one output pixel, from staged planes.
//   win: top-left of the window in plane 0 of the staged buffer; wdw: per-channel [KS*KS taps, scale, shift, pad]
template <int KS, int S, bool RELU>
__device__ __forceinline__ void dw8p(const float* __restrict__ xk, int RS, int WS, const float* __restrict__ wk, bool valid, float (&a)[8]);
template <int KS, int S, bool RELU>
__device__ __forceinline__ void dw8(const float* __restrict__ win, int RS, int WS, const float* __restrict__ wdw, int k0, bool valid,
                                    float (&a)[8]) {
    dw8p<KS, S, RELU>(win + k0 * RS, RS, WS, wdw + k0 * (KS == 3 ? 12 : 28), valid, a);
}
// same, given the window in the first of the 8 planes (xk, plane stride RS) and the first channel's weight row (wk)
template <int KS, int S, bool RELU>
__device__ __forceinline__ void dw8p(const float* __restrict__ xk, int RS, int WS, const float* __restrict__ wk, bool valid, float (&a)[8]) {
    constexpr int R = KS == 3 ? 12 : 28;
    if (!valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = 0.f;
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float w[R];
#pragma unroll
        for (int t = 0; t < R / 4; ++t) {
            const float4 w4 = *reinterpret_cast<const float4*>(wk + 4 * t);
            w[4 * t] = w4.x; w[4 * t + 1] = w4.y; w[4 * t + 2] = w4.z; w[4 * t + 3] = w4.w;
        }
        // one independent partial sum per kernel row (breaks the 9- / 25-long dependent FMA chain), then a short add tree
        float pr[KS];
        const float* row = xk;
#pragma unroll
        for (int dy = 0; dy < KS; ++dy) {
            pr[dy] = w[dy * KS] * row[0];
#pragma unroll
            for (int dx = 1; dx < KS; ++dx) pr[dy] = fmaf(w[dy * KS + dx], row[dx], pr[dy]);
            row += WS;
        }
        float d = KS == 3 ? (pr[0] + pr[1]) + pr[2] : ((pr[0] + pr[1]) + (pr[2] + pr[3])) + pr[KS - 1];
        d = fmaf(d, w[KS * KS], w[KS * KS + 1]);
        if (RELU) d = fmaxf(d, 0.f);
        a[j] = d;
        xk += RS;
        wk += R;
    }
}

// ===================================================================================================
// tc_pw_kernel: pointwise over planes.  KA channels read through (y>>SHA, x>>SHA) from A, then KB channels
// at (y,x) from B (K = KA+KB must be a multiple of 48).  One tile = 128 consecutive pixels of the flattened
// (image, pixel) index space.
// ===================================================================================================
struct PwArgs {
    Planes A, B, out;
    ChanTab ta, tb, tout;
    const float* wpack;     // tc pack: Bhi | Blo | scale | shift
    int N;                  // images
    int nout;               // real output channels
};

template <int KA, int KB, int SHA, int NP, int G, bool RELU, int NB = 2>
__global__ void __launch_bounds__(G * 128, (G * (16 * NB + NP) > kTmemCols) ? 1 : 2)
tc_pw_kernel(const __grid_constant__ PwArgs p) {
    pdl_trigger();
    constexpr int KP = KA + KB;
    constexpr int PF = 48;                                   // channels prefetched per batch of global loads
    static_assert(KP % PF == 0 && NP % 16 == 0, "shape");
    constexpr int COLS = 16 * NB + NP;                       // NB-deep A ring + accumulator
    constexpr int TOT = (G * COLS > kTmemCols) ? 512 : kTmemCols;   // 4 groups: the CTA owns the SM's whole TMEM
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) Pipe pipes[G];
    __shared__ uint32_t tmem_slot;
    float* sB = smem;
    constexpr int WFL = 2 * NP * KP + 2 * NP;
    copy_f4(sB, p.wpack, WFL, G * 128);
    publish_smem();
    Grp g = cta_setup<G, COLS, TOT>(pipes, &tmem_slot);
    pdl_wait();                                            // predecessor's activations are complete and visible from here on
    const uint32_t b_hi = smem_u32(sB), b_lo = smem_u32(sB + NP * KP);
    const float* scale = sB + 2 * NP * KP;
    const float* shift = scale + NP;
    const int HW = p.out.H * p.out.W, W = p.out.W;
    const long long total = (long long)p.N * HW;
    const int ntiles = (int)((total + 127) / 128);
    const int grp = threadIdx.x >> 7;
    const unsigned sCa = (unsigned)p.A.sC, sCb = (unsigned)p.B.sC, sCo = (unsigned)p.out.sC;
    for (int tile = blockIdx.x * G + grp; tile < ntiles; tile += gridDim.x * G) {
        const long long pos = (long long)tile * 128 + g.gtid;
        const bool valid = pos < total;
        const int n = valid ? (int)(pos / HW) : 0;
        const int px = valid ? (int)(pos - (long long)n * HW) : 0;
        const int y = px / W, x = px - y * W;
        const float* baseA = p.A.base + (long long)n * p.A.sN + p.A.org + (y >> SHA) * p.A.Ws + (x >> SHA);
        const float* baseB = p.B.base + (long long)n * p.B.sN + p.B.org + y * p.B.Ws + x;
        float* obase = p.out.base + (long long)n * p.out.sN + p.out.org + y * p.out.Ws + x;
#pragma unroll 1
        for (int kb = 0; kb < KP; kb += PF) {
            float v[PF];
#pragma unroll
            for (int j = 0; j < PF; ++j) {
                const int k = kb + j;
                v[j] = 0.f;
                if (valid) v[j] = (k < KA) ? __ldg(baseA + p.ta.c[k] * sCa) : __ldg(baseB + p.tb.c[KB ? k - KA : 0] * sCb);
            }
#pragma unroll
            for (int c = 0; c < PF / 8; ++c) {
                float a[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = v[c * 8 + j];
                put_chunk<KP, NP, NB>(g, a, kb / 8 + c, b_hi, b_lo);
            }
        }
        get_tile<NP, NB>(g, [&](int n0, float (&d)[16]) {
            if (valid) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int nn = n0 + j;
                    if (nn < p.nout) {
                        float r = fmaf(d[j], scale[nn], shift[nn]);
                        if (RELU) r = fmaxf(r, 0.f);
                        obase[p.tout.c[nn] * sCo] = r;
                    }
                }
            }
        });
    }
    cta_teardown<TOT>(&tmem_slot);
}

// ===================================================================================================
// tc_dwpw_kernel: DW(KSxKS, stride S)+BN(+ReLU) -> PW(K->NP)+BN(+ReLU) -> planes, or (DENSE) -> dense NCHW prediction
// tensors (the heads' second half: its pointwise, BN and the shared output conv are one folded matrix, see plan.cu).
// A work item is (image group, band of TR output rows, branch); the band (+halo) of the K source planes of each image
// is staged in shared memory by bulk copies.  Used by the stage-4 blocks and as the heads' fallback for maps of more
// than 512 pixels (tc_head_kernel below is the fast path).
// ===================================================================================================
struct DwPwArgs {
    Planes in[2], out[2];      // per branch
    ChanTab tin[2], tout[2];
    const float* wdw[2];       // DW pack per branch
    const float* wpw[2];       // tc pack per branch
    float* dstA[2]; float* dstB[2]; int split[2]; int M[2];   // DENSE: channels [0,split) -> dstA, [split,M) -> dstB
    int N, TR, bandsPerImg, nbranch, nout;
    int imgs;                  // images per work item (> 1 only when a band is a whole image)
};

// epilogue of one 16-column block of a pixel's output row
template <bool RELU_OUT, bool DENSE, bool IDENT = false>
struct RowSink {
    const float* scale; const float* shift;
    float* obase; const unsigned short* tout; unsigned sCo; int nout;          // planes
    float* dA; float* dB; int split, M; long long n, HW, opix;                  // dense
    bool valid;
    __device__ __forceinline__ void operator()(int n0, float (&d)[16]) const {
        if (!valid) return;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int m = n0 + j;
            float v = fmaf(d[j], scale[m], shift[m]);
            if (RELU_OUT) v = fmaxf(v, 0.f);
            if (!DENSE) { if (m < nout) obase[(IDENT ? (unsigned)m : (unsigned)tout[m]) * sCo] = v; }
            else if (m < split) dA[(n * split + m) * HW + opix] = v;
            else if (m < M) dB[(n * (M - split) + (m - split)) * HW + opix] = v;
        }
    }
};

template <int K, int NP, int G, int KS, int S, bool RELU_DW, bool RELU_OUT, bool DENSE>
__global__ void __launch_bounds__(G * 128, 1)
tc_dwpw_kernel(const __grid_constant__ DwPwArgs p) {
    pdl_trigger();
    constexpr int KP = K;
    static_assert(KP % 8 == 0 && NP % 16 == 0 && K <= G * 128, "shape");
    constexpr int COLS = kACols + NP;
    constexpr int PADK = KS / 2;
    constexpr int DWR = KS == 3 ? 12 : 28;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) Pipe pipes[G];
    __shared__ __align__(8) uint64_t xbar;
    __shared__ uint32_t tmem_slot;
    constexpr int WFL = 2 * NP * KP + 2 * NP;
    float* sB = smem;                          // pw pack
    float* sDW = sB + WFL;                     // dw pack
    float* X = sDW + K * DWR;                  // staged planes
    const int Hout = DENSE ? p.in[0].H : p.out[0].H, Wout = DENSE ? p.in[0].W : p.out[0].W;
    const int WS = p.in[0].Ws;
    const int RS1 = (S * (p.TR - 1) + KS) * WS;           // one image's band of one plane
    const int RS = RS1 * p.imgs;                           // plane stride of the staged buffer
    const int coff = p.in[0].pad - PADK;                   // frame column of the window's left edge for ox = 0
    if (threadIdx.x == 0) { mbar_init(&xbar, 1); fence_mbar_init(); }
    int loaded_branch = (int)(blockIdx.x % p.nbranch);     // the first item's weights are part of the prologue
    copy_f4(sB, p.wpw[loaded_branch], WFL, G * 128);
    copy_f4(sDW, p.wdw[loaded_branch], K * DWR, G * 128);
    publish_smem();
    Grp g = cta_setup<G, COLS>(pipes, &tmem_slot);
    pdl_wait();                                            // predecessor's activations are complete and visible from here on
    const uint32_t b_hi = smem_u32(sB), b_lo = smem_u32(sB + NP * KP);
    const float* scale = sB + 2 * NP * KP;
    const float* shift = scale + NP;
    const int ngroups = (p.N + p.imgs - 1) / p.imgs;       // image groups
    const int items = ngroups * p.bandsPerImg * p.nbranch;
    const int grp = threadIdx.x >> 7;
    uint32_t xparity = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int br = item % p.nbranch;
        const int rem = item / p.nbranch;
        const int ig = rem / p.bandsPerImg;
        const int r0 = (rem - ig * p.bandsPerImg) * p.TR;
        const int rows = min(p.TR, Hout - r0);
        const int n0img = ig * p.imgs;
        const int nimg = min(p.imgs, p.N - n0img);
        const int nrows = S * (rows - 1) + KS;
        __syncthreads();                                   // previous item done with X / weights
        publish_smem();                                    // order those generic reads before the async-proxy writes below
        if (threadIdx.x == 0) mbar_expect_tx(&xbar, (uint32_t)(K * nimg * nrows * WS * sizeof(float)));
        for (int i = 0; i < nimg; ++i)
            stage_bulk<K>(X + i * RS1, RS, p.in[br], p.tin[br], n0img + i, S * r0 + p.in[br].pad - PADK, nrows, &xbar, 0);
        if (br != loaded_branch) {
            copy_f4(sB, p.wpw[br], WFL, G * 128);
            copy_f4(sDW, p.wdw[br], K * DWR, G * 128);
            loaded_branch = br;
            publish_smem();
        }
        __syncthreads();
        mbar_wait(&xbar, xparity);
        xparity ^= 1u;
        const int ppi = rows * Wout;                        // pixels per image in this band
        const int npix = ppi * nimg;
        for (int tile = grp; tile * 128 < npix; tile += G) {
            const int q = tile * 128 + g.gtid;
            const bool valid = q < npix;
            const int im = valid ? q / ppi : 0;
            const int qi = valid ? q - im * ppi : 0;
            const int n = n0img + im;
            const int orow = qi / Wout, ox = qi - orow * Wout;
            const float* win = X + im * RS1 + (S * orow) * WS + S * ox + coff;
            RowSink<RELU_OUT, DENSE> sink;
            sink.scale = scale; sink.shift = shift; sink.valid = valid;
            if (!DENSE) {
                sink.obase = p.out[br].base + (long long)n * p.out[br].sN + p.out[br].org + (r0 + orow) * p.out[br].Ws + ox;
                sink.tout = p.tout[br].c; sink.sCo = (unsigned)p.out[br].sC; sink.nout = p.nout;
            } else {
                sink.dA = p.dstA[br]; sink.dB = p.dstB[br]; sink.split = p.split[br]; sink.M = p.M[br];
                sink.n = n; sink.HW = (long long)Hout * Wout; sink.opix = (long long)(r0 + orow) * Wout + ox;
            }
            pw_tile_rolled<KP, NP>(g, b_hi, b_lo,
                [&](int k0, float (&a)[8]) { dw8<KS, S, RELU_DW>(win, RS, WS, sDW, k0, valid, a); }, sink);
        }
    }
    cta_teardown(&tmem_slot);
}

// ===================================================================================================
// tc_head_kernel: the detection heads' DW5x5+BN+ReLU -> PW, channel-streamed (reference fpn.py DWConvblock, detector.py
// output convs).  One work item = (group of `imgs` whole images, branch) with at most G*128 pixels, one 128-pixel tile
// per warpgroup, so the CTA never keeps more than 8 input channels of the item in shared memory: a dedicated producer
// warp streams the framed planes of channels [8c, 8c+8) through a ring of kHeadBufs buffers with one bulk copy per
// plane, the G warpgroups run the stencil on chunk c while chunks c+1.. are in flight and the tensor core contracts
// chunk c-1.  Compared with tc_dwpw_kernel (whole 72-channel band resident) this needs ~1/9 of the staging memory, which
// buys 16 compute warps per SM instead of 8 and whole-image tiles (no band halo re-reads, 95 % tile fill).
// ===================================================================================================
struct HeadArgs {
    Planes in[2], out[2];      // per branch (ident channel tables)
    const float* wdw[2];
    const float* wpw[2];
    float* dstA[2]; float* dstB[2]; int split[2]; int M[2];   // DENSE
    int N, imgs, nout;
    int band_rows, bands;      // tc_head2w_kernel on large maps: an item is one band of `band_rows` rows of one image (0: whole images)
    // tc_head2w_kernel, whole-image items: pixel pair of thread q of the item as (image << 16 | row << 8 | pair column), 0xFFFFFFFF:
    // idle lane.  Raster order makes every half-warp of the stencil's LDS.64 window loads 2-way bank conflicted (a half-warp spans
    // two framed rows whose bank ranges overlap; ncu: 9.2 M conflict wavefronts per launch); build_lanemap() deals the pairs so that
    // the 16 lanes of a half-warp sit in 16 different 8-byte banks.
    int use_map;
    int st2_ok;                // tc_head2w_kernel: a pixel pair may leave as one 8-byte store (even W, 8-byte aligned destinations)
    unsigned int lanemap[256];
};
constexpr int kHeadBufs = 4;

template <int K, int NP, int G, bool DENSE>
__global__ void __launch_bounds__(G * 128 + 32, 1)
tc_head_kernel(const __grid_constant__ HeadArgs p) {
    pdl_trigger();
    constexpr int NCH = K / 8, NB = kHeadBufs, DWR = 28, COLS = 128, TOT = 512;
    static_assert(K % 8 == 0 && NP % 16 == 0 && kACols + NP <= COLS && G * COLS <= TOT, "shape");
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) Pipe pipes[G + 1];             // +1: the producer warp's (unused) slot
    __shared__ __align__(8) uint64_t fullb[NB], freeb[NB];
    __shared__ uint32_t tmem_slot;
    constexpr int WFL = 2 * NP * K + 2 * NP;
    float* sB = smem;
    float* sDW = sB + WFL;
    float* X = sDW + K * DWR;
    const int H = p.in[0].H, W = p.in[0].W, WS = p.in[0].Ws;
    const int PS = (H + 4) * WS;                            // one whole framed plane (frame of 2)
    const int CS = PS * p.imgs;                             // channel stride inside a ring buffer
    const int BUF = 8 * CS;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NB; ++i) { mbar_init(&fullb[i], 1); mbar_init(&freeb[i], G * 4); }
        fence_mbar_init();
    }
    const int ngroups = (p.N + p.imgs - 1) / p.imgs;
    const int items = 2 * ngroups;                          // branch-major: item t -> (branch t / ngroups, group t % ngroups)
    const int first_branch = (int)blockIdx.x / ngroups;     // < 2: the grid never exceeds `items`
    if (threadIdx.x < G * 128) {                            // the first item's weights are part of the prologue
        copy_f4(sB, p.wpw[first_branch], WFL, G * 128);
        copy_f4(sDW, p.wdw[first_branch], K * DWR, G * 128);
        publish_smem();
    }
    Grp g = cta_setup<G, COLS, TOT>(pipes, &tmem_slot);
    pdl_wait();                                            // predecessor's activations are complete and visible from here on
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t it = 0;                                        // chunks produced / consumed so far (same sequence on both sides)
    if (warp == G * 4) {
        // ---------------- producer warp ------------------------------------------------------------------------
        for (int t = blockIdx.x; t < items; t += gridDim.x) {
            const int br = t / ngroups, n0 = (t - br * ngroups) * p.imgs;
            const int nimg = min(p.imgs, p.N - n0);
            for (int c = 0; c < NCH; ++c, ++it) {
                const uint32_t buf = it % NB, use = it / NB;
                if (use > 0) mbar_wait(&freeb[buf], (use - 1) & 1u);     // all 4G compute warps are done reading this buffer
                publish_smem();
                if (lane == 0) mbar_expect_tx(&fullb[buf], (uint32_t)(8 * nimg * PS * sizeof(float)));
                __syncwarp();
                for (int j = lane; j < 8 * nimg; j += 32) {
                    const int ch = j & 7, i = j >> 3;
                    bulk_g2s(X + (size_t)buf * BUF + ch * CS + i * PS, plane_ptr(p.in[br], n0 + i, c * 8 + ch),
                             (uint32_t)(PS * sizeof(float)), &fullb[buf]);
                }
            }
        }
    } else {
        // ---------------- compute warpgroups ---------------------------------------------------------------------
        const uint32_t b_hi = smem_u32(sB), b_lo = smem_u32(sB + NP * K);
        const float* scale = sB + 2 * NP * K;
        const float* shift = scale + NP;
        const int grp = threadIdx.x >> 7;
        const int HW = H * W;
        int loaded_branch = first_branch;
        for (int t = blockIdx.x; t < items; t += gridDim.x) {
            const int br = t / ngroups, n0 = (t - br * ngroups) * p.imgs;
            const int nimg = min(p.imgs, p.N - n0);
            if (br != loaded_branch) {
                group_bar(1, G * 128);                      // every compute warp is done with the previous branch's weights
                copy_f4(sB, p.wpw[br], WFL, G * 128);
                copy_f4(sDW, p.wdw[br], K * DWR, G * 128);
                publish_smem();
                group_bar(1, G * 128);
                loaded_branch = br;
            }
            const int q = grp * 128 + g.gtid;
            const bool valid = q < HW * nimg;
            const int im = valid ? q / HW : 0;
            const int qi = valid ? q - im * HW : 0;
            const int oy = qi / W, ox = qi - oy * W;
            const int woff = im * PS + oy * WS + ox;        // window's top-left in the framed plane (frame 2 = the 5x5 halo)
            RowSink<false, DENSE, true> sink;
            sink.scale = scale; sink.shift = shift; sink.valid = valid;
            if (!DENSE) {
                sink.obase = p.out[br].base + (long long)(n0 + im) * p.out[br].sN + p.out[br].org + oy * p.out[br].Ws + ox;
                sink.tout = nullptr; sink.sCo = (unsigned)p.out[br].sC; sink.nout = p.nout;
            } else {
                sink.dA = p.dstA[br]; sink.dB = p.dstB[br]; sink.split = p.split[br]; sink.M = p.M[br];
                sink.n = n0 + im; sink.HW = HW; sink.opix = qi;
            }
#pragma unroll 1
            for (int c = 0; c < NCH; ++c, ++it) {
                const uint32_t buf = it % NB;
                mbar_wait(&fullb[buf], (it / NB) & 1u);
                float a[8];
                dw8p<5, 1, true>(X + (size_t)buf * BUF + woff, CS, WS, sDW + c * 8 * DWR, valid, a);
                __syncwarp();
                if (lane == 0) mbar_arrive(&freeb[buf]);
                put_chunk<K, NP>(g, a, c, b_hi, b_lo);
            }
            get_tile<NP>(g, sink, DENSE ? p.M[br] : p.nout);
        }
    }
    cta_teardown<TOT>(&tmem_slot);
}

// ===================================================================================================
// tc_dws2c_kernel: stride-2 depthwise 3x3 + BN -> pointwise + BN + ReLU on SMALL output maps, channel-streamed like the heads
// (reference shufflenetv2.py:34-44: branch_proj and the tail of branch_main of a stride-2 block; K = 96: stage 4).
// tc_dwpw_kernel<96,...> staged whole 96-channel bands for two warpgroups: ncu showed 13 % issue slots, 6.9 barrier stalls per issue,
// 127 us for 71 MB.  Here a work item is (a group of `imgs` whole images, branch): one 128-pixel tile per warpgroup (an 11x11 map
// is one tile), a producer warp streams the framed input planes four channels at a time through a ring of kDwsBufs slots (one
// bulk copy per plane) and the stencil of an 8-channel chunk reads two slots.
// ===================================================================================================
struct Dws2Args {
    Planes in[2], out[2];
    ChanTab tin[2], tout[2];
    const float* wdw[2];
    const float* wpw[2];
    int N, imgs, nbranch, nout;
    int pair_rows;             // 1: window rows as LDS.64 + shuffle (frames of odd width: every window starts on an even float)
};
constexpr int kDwsBufs = 3;

// Stride-2 rows without the 2-way bank conflict of scalar loads (consecutive lanes sit two floats apart and the frame pitch puts
// every lane on an even bank: ncu on tc_dws2c_kernel<96,96,4>: 4.8 MIO-throttle stalls per issue, 4.7 M conflict wavefronts): the
// first two taps of a window row come as ONE 8-byte load and the third is the right neighbour's first tap (one shuffle); lanes
// whose right neighbour is another row / image / warp (`nb_ok` false) load it.  Every lane of the warp must call this (shuffles);
// lanes without a pixel pass the window of pixel 0 and `valid` = false.  Same arithmetic and association as dw3n: bit-identical.
template <int NC>
__device__ __forceinline__ void dw3n_s2(const float* __restrict__ xk, int RS, int WS, const float* __restrict__ wk, bool valid, bool nb_ok, float* a) {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const float4 wa = *reinterpret_cast<const float4*>(wk);
        const float4 wb = *reinterpret_cast<const float4*>(wk + 4);
        const float4 wc = *reinterpret_cast<const float4*>(wk + 8);
        float2 v[3];
        float c2[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            v[r] = *reinterpret_cast<const float2*>(xk + r * WS);
            c2[r] = __shfl_down_sync(0xffffffffu, v[r].x, 1);
            if (!nb_ok) c2[r] = xk[r * WS + 2];
        }
        float p0 = wa.x * v[0].x; p0 = fmaf(wa.y, v[0].y, p0); p0 = fmaf(wa.z, c2[0], p0);
        float p1 = wa.w * v[1].x; p1 = fmaf(wb.x, v[1].y, p1); p1 = fmaf(wb.y, c2[1], p1);
        float p2 = wb.z * v[2].x; p2 = fmaf(wb.w, v[2].y, p2); p2 = fmaf(wc.x, c2[2], p2);
        const float d = (p0 + p1) + p2;
        a[j] = valid ? fmaf(d, wc.y, wc.z) : 0.f;
        xk += RS;
        wk += 12;
    }
}

// depthwise 3x3 + BN of NC consecutive channels of one output pixel (window top-left xk in the first plane, plane stride RS)
template <int NC>
__device__ __forceinline__ void dw3n(const float* __restrict__ xk, int RS, int WS, const float* __restrict__ wk, bool valid, float* a) {
    if (!valid) {
#pragma unroll
        for (int j = 0; j < NC; ++j) a[j] = 0.f;
        return;
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const float4 wa = *reinterpret_cast<const float4*>(wk);
        const float4 wb = *reinterpret_cast<const float4*>(wk + 4);
        const float4 wc = *reinterpret_cast<const float4*>(wk + 8);
        const float* r0 = xk; const float* r1 = xk + WS; const float* r2 = xk + 2 * WS;
        float p0 = wa.x * r0[0]; p0 = fmaf(wa.y, r0[1], p0); p0 = fmaf(wa.z, r0[2], p0);
        float p1 = wa.w * r1[0]; p1 = fmaf(wb.x, r1[1], p1); p1 = fmaf(wb.y, r1[2], p1);
        float p2 = wb.z * r2[0]; p2 = fmaf(wb.w, r2[1], p2); p2 = fmaf(wc.x, r2[2], p2);
        const float d = (p0 + p1) + p2;                     // same association as dw8p<3,...>
        a[j] = fmaf(d, wc.y, wc.z);
        xk += RS;
        wk += 12;
    }
}

template <int K, int NP, int G>
__global__ void __launch_bounds__(G * 128 + 32, 1)
tc_dws2c_kernel(const __grid_constant__ Dws2Args p) {
    pdl_trigger();
    constexpr int NCH = K / 8, NB = kDwsBufs, DWR = 12, COLS = 128, TOT = 512;
    static_assert(K % 8 == 0 && NP % 16 == 0 && kACols + NP <= COLS && G * COLS <= TOT, "shape");
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) Pipe pipes[G + 1];
    __shared__ __align__(8) uint64_t fullb[NB], freeb[NB];
    __shared__ uint32_t tmem_slot;
    constexpr int WFL = 2 * NP * K + 2 * NP;
    float* sB = smem;
    float* sDW = sB + WFL;
    float* X = sDW + K * DWR;
    const int WS = p.in[0].Ws, pad = p.in[0].pad;
    const int PS = (p.in[0].H + 2 * pad) * WS;              // one whole framed input plane
    const int CS = PS * p.imgs;                             // channel stride inside a ring slot
    const int SLOT = 4 * CS;
    const int Hout = p.out[0].H, Wout = p.out[0].W, HWo = Hout * Wout;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NB; ++i) { mbar_init(&fullb[i], 1); mbar_init(&freeb[i], G * 4); }
        fence_mbar_init();
    }
    const int ngroups = (p.N + p.imgs - 1) / p.imgs;
    const int items = p.nbranch * ngroups;                  // branch-major
    const int first_branch = (int)blockIdx.x / ngroups;
    if (threadIdx.x < G * 128) {
        copy_f4(sB, p.wpw[first_branch], WFL, G * 128);
        copy_f4(sDW, p.wdw[first_branch], K * DWR, G * 128);
        publish_smem();
    }
    Grp g = cta_setup<G, COLS, TOT>(pipes, &tmem_slot);
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t it = 0;                                        // ring slots produced / consumed so far
    if (warp == G * 4) {
        for (int t = blockIdx.x; t < items; t += gridDim.x) {
            const int br = t / ngroups, n0 = (t - br * ngroups) * p.imgs;
            const int nimg = min(p.imgs, p.N - n0);
            for (int c = 0; c < 2 * NCH; ++c, ++it) {       // slot c of the item: channels [4c, 4c + 4)
                const uint32_t buf = it % NB, use = it / NB;
                if (use > 0) mbar_wait(&freeb[buf], (use - 1) & 1u);
                publish_smem();
                if (lane == 0) mbar_expect_tx(&fullb[buf], (uint32_t)(4 * nimg * PS * sizeof(float)));
                __syncwarp();
                for (int j = lane; j < 4 * nimg; j += 32) {
                    const int ch = j & 3, i = j >> 2;
                    bulk_g2s(X + (size_t)buf * SLOT + ch * CS + i * PS, plane_ptr(p.in[br], n0 + i, p.tin[br].c[c * 4 + ch]),
                             (uint32_t)(PS * sizeof(float)), &fullb[buf]);
                }
            }
        }
    } else {
        const uint32_t b_hi = smem_u32(sB), b_lo = smem_u32(sB + NP * K);
        const float* scale = sB + 2 * NP * K;
        const float* shift = scale + NP;
        const int grp = threadIdx.x >> 7;
        int loaded_branch = first_branch;
        for (int t = blockIdx.x; t < items; t += gridDim.x) {
            const int br = t / ngroups, n0 = (t - br * ngroups) * p.imgs;
            const int nimg = min(p.imgs, p.N - n0);
            if (br != loaded_branch) {
                group_bar(1, G * 128);
                copy_f4(sB, p.wpw[br], WFL, G * 128);
                copy_f4(sDW, p.wdw[br], K * DWR, G * 128);
                publish_smem();
                group_bar(1, G * 128);
                loaded_branch = br;
            }
            const int q = grp * 128 + g.gtid;
            const bool valid = q < HWo * nimg;
            const int im = valid ? q / HWo : 0;
            const int qi = valid ? q - im * HWo : 0;
            const int oy = qi / Wout, ox = qi - oy * Wout;
            const int woff = im * PS + (2 * oy + pad - 1) * WS + 2 * ox + pad - 1;      // window's top-left in the framed plane
            const bool nb_ok = valid && ox + 1 < Wout && lane != 31;                    // the next lane holds the pixel to the right
            RowSink<true, false, false> sink;
            sink.scale = scale; sink.shift = shift; sink.valid = valid;
            sink.obase = p.out[br].base + (long long)(n0 + im) * p.out[br].sN + p.out[br].org + oy * p.out[br].Ws + ox;
            sink.tout = p.tout[br].c; sink.sCo = (unsigned)p.out[br].sC; sink.nout = p.nout;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                float a[8];
#pragma unroll
                for (int h = 0; h < 2; ++h, ++it) {
                    const uint32_t buf = it % NB;
                    mbar_wait(&fullb[buf], (it / NB) & 1u);
                    if (p.pair_rows) dw3n_s2<4>(X + (size_t)buf * SLOT + woff, CS, WS, sDW + (c * 8 + 4 * h) * DWR, valid, nb_ok, a + 4 * h);
                    else dw3n<4>(X + (size_t)buf * SLOT + woff, CS, WS, sDW + (c * 8 + 4 * h) * DWR, valid, a + 4 * h);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&freeb[buf]);
                }
                put_chunk<K, NP>(g, a, c, b_hi, b_lo);
            }
            get_tile<NP>(g, sink, p.nout);
        }
    }
    cta_teardown<TOT>(&tmem_slot);
}

// ===================================================================================================
// tc_head2_kernel: tc_head_kernel with TWO pixels per thread.  The 5x5 stencil is bound by shared-memory wavefronts
// (ncu: 48 % of the stalls are MIO throttle, 25 LDS per pixel-channel), so a thread takes a horizontally adjacent pixel
// pair (x even): the six window columns of a row arrive as three LDS.64 and feed both pixels, the weights are loaded
// once for the pair -> 18.5 instead of 32 wavefronts per 32 pixel-channels and 1.5x fewer instructions.  A thread can
// only write its own TMEM lane, so the two pixels of a pair are the same row of two different M=128 tiles: a
// warpgroup owns two accumulators (columns [64, 64+NP) and [64+NP, 64+2NP)) and the A ring holds both tiles' chunks
// ([hi8|lo8] of tile 0, [hi8|lo8] of tile 1, twice).  2 warpgroups x 256 columns fill the SM's TMEM.
// ===================================================================================================
template <bool RELU>
__device__ __forceinline__ void dw5_pair8(const float* __restrict__ xk, int RS, int WS, const float* __restrict__ wk, bool valid0, bool valid1,
                                          float (&a0)[8], float (&a1)[8]) {
    if (!valid0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float w[28];
#pragma unroll
        for (int t = 0; t < 7; ++t) {
            const float4 w4 = *reinterpret_cast<const float4*>(wk + 4 * t);
            w[4 * t] = w4.x; w[4 * t + 1] = w4.y; w[4 * t + 2] = w4.z; w[4 * t + 3] = w4.w;
        }
        float p0[5], p1[5];
        const float* row = xk;
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) {
            const float2 v01 = *reinterpret_cast<const float2*>(row);
            const float2 v23 = *reinterpret_cast<const float2*>(row + 2);
            const float2 v45 = *reinterpret_cast<const float2*>(row + 4);
            const float v[6] = {v01.x, v01.y, v23.x, v23.y, v45.x, v45.y};
            p0[dy] = w[dy * 5] * v[0];
            p1[dy] = w[dy * 5] * v[1];
#pragma unroll
            for (int dx = 1; dx < 5; ++dx) {
                p0[dy] = fmaf(w[dy * 5 + dx], v[dx], p0[dy]);
                p1[dy] = fmaf(w[dy * 5 + dx], v[dx + 1], p1[dy]);
            }
            row += WS;
        }
        float d0 = ((p0[0] + p0[1]) + (p0[2] + p0[3])) + p0[4];      // same association as dw8p
        float d1 = ((p1[0] + p1[1]) + (p1[2] + p1[3])) + p1[4];
        d0 = fmaf(d0, w[25], w[26]);
        d1 = fmaf(d1, w[25], w[26]);
        if (RELU) { d0 = fmaxf(d0, 0.f); d1 = fmaxf(d1, 0.f); }
        a0[j] = d0;
        a1[j] = valid1 ? d1 : 0.f;
        xk += RS;
        wk += 28;
    }
}

constexpr int kA2Cols = 64;             // 2 buffers x 2 tiles x (8 hi + 8 lo)

template <int KP, int NP>
__device__ __forceinline__ void put_chunk2(Grp& g, const float (&a0)[8], const float (&a1)[8], int c, uint32_t b_hi, uint32_t b_lo) {
    const uint32_t buf = g.chunk & 1u, use = g.chunk >> 1;
    if (use > 0) mbar_wait(&g.pipe->empty[buf], (use - 1) & 1u);
    fence_after_sync();
    store_a8(g.tlane + buf * 32, 8, a0);
    store_a8(g.tlane + buf * 32 + 16, 8, a1);
    wait_st();
    fence_before_sync();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const uint32_t old = atom_add_acq_rel(&g.pipe->arrivals[buf], 1u);
        if ((old & 3u) == 3u) {
            fence_after_sync();
            constexpr uint32_t idesc = make_idesc_tf32(128, NP);
            constexpr uint32_t LBO = 128, SBO = (KP / 4) * 128;
            const uint64_t bh = make_b_desc(b_hi + c * 256, LBO, SBO);
            const uint64_t bl = make_b_desc(b_lo + c * 256, LBO, SBO);
#pragma unroll
            for (int tile = 0; tile < 2; ++tile) {
                const uint32_t a_hi = g.tcol + buf * 32 + tile * 16, a_lo = a_hi + 8, d = g.tcol + kA2Cols + tile * NP;
                mma_tf32_ts(d, a_lo, bh, idesc, c > 0 ? 1u : 0u);
                mma_tf32_ts(d, a_hi, bl, idesc, 1u);
                mma_tf32_ts(d, a_hi, bh, idesc, 1u);
            }
            mma_commit(&g.pipe->empty[buf]);
            if (c == KP / 8 - 1) mma_commit(&g.pipe->dfull);
        }
    }
    __syncwarp();
    ++g.chunk;
}
template <int NP, class Epi>
__device__ __forceinline__ void get_tiles2(Grp& g, const Epi& epi0, const Epi& epi1, int ncols) {
    mbar_wait(&g.pipe->dfull, g.dparity);
    g.dparity ^= 1u;
    fence_after_sync();
#pragma unroll 1
    for (int n0 = 0; n0 < ncols; n0 += 16) {                // rolled: the dense epilogue is long and identical per block
        float d[16];
        tmem_ld16(g.tlane + kA2Cols + n0, d);
        wait_ld();
        epi0(n0, d);
    }
#pragma unroll 1
    for (int n0 = 0; n0 < ncols; n0 += 16) {
        float d[16];
        tmem_ld16(g.tlane + kA2Cols + NP + n0, d);
        wait_ld();
        epi1(n0, d);
    }
}

template <int K, int NP, bool DENSE>
__global__ void __launch_bounds__(2 * 128 + 32, 1)
tc_head2_kernel(const __grid_constant__ HeadArgs p) {
    pdl_trigger();
    constexpr int G = 2, NCH = K / 8, NB = kHeadBufs, DWR = 28, COLS = 256, TOT = 512;
    static_assert(K % 8 == 0 && NP % 16 == 0 && kA2Cols + 2 * NP <= COLS, "shape");
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) Pipe pipes[G + 1];
    __shared__ __align__(8) uint64_t fullb[NB], freeb[NB];
    __shared__ uint32_t tmem_slot;
    constexpr int WFL = 2 * NP * K + 2 * NP;
    float* sB = smem;
    float* sDW = sB + WFL;
    float* X = sDW + K * DWR;
    const int H = p.in[0].H, W = p.in[0].W, WS = p.in[0].Ws;
    const int PS = (H + 4) * WS;
    const int CS = PS * p.imgs;
    const int BUF = 8 * CS;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NB; ++i) { mbar_init(&fullb[i], 1); mbar_init(&freeb[i], G * 4); }
        fence_mbar_init();
    }
    const int ngroups = (p.N + p.imgs - 1) / p.imgs;
    const int items = 2 * ngroups;
    const int first_branch = (int)blockIdx.x / ngroups;
    if (threadIdx.x < G * 128) {
        copy_f4(sB, p.wpw[first_branch], WFL, G * 128);
        copy_f4(sDW, p.wdw[first_branch], K * DWR, G * 128);
        publish_smem();
    }
    Grp g = cta_setup<G, COLS, TOT>(pipes, &tmem_slot);
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t it = 0;
    if (warp == G * 4) {
        for (int t = blockIdx.x; t < items; t += gridDim.x) {
            const int br = t / ngroups, n0 = (t - br * ngroups) * p.imgs;
            const int nimg = min(p.imgs, p.N - n0);
            for (int c = 0; c < NCH; ++c, ++it) {
                const uint32_t buf = it % NB, use = it / NB;
                if (use > 0) mbar_wait(&freeb[buf], (use - 1) & 1u);
                publish_smem();
                if (lane == 0) mbar_expect_tx(&fullb[buf], (uint32_t)(8 * nimg * PS * sizeof(float)));
                __syncwarp();
                for (int j = lane; j < 8 * nimg; j += 32) {
                    const int ch = j & 7, i = j >> 3;
                    bulk_g2s(X + (size_t)buf * BUF + ch * CS + i * PS, plane_ptr(p.in[br], n0 + i, c * 8 + ch),
                             (uint32_t)(PS * sizeof(float)), &fullb[buf]);
                }
            }
        }
    } else {
        const uint32_t b_hi = smem_u32(sB), b_lo = smem_u32(sB + NP * K);
        const float* scale = sB + 2 * NP * K;
        const float* shift = scale + NP;
        const int grp = threadIdx.x >> 7;
        const int HW = H * W;
        const int Wp = (W + 1) >> 1, PPI = H * Wp;          // pixel pairs per row / per image
        int loaded_branch = first_branch;
        for (int t = blockIdx.x; t < items; t += gridDim.x) {
            const int br = t / ngroups, n0 = (t - br * ngroups) * p.imgs;
            const int nimg = min(p.imgs, p.N - n0);
            if (br != loaded_branch) {
                group_bar(1, G * 128);
                copy_f4(sB, p.wpw[br], WFL, G * 128);
                copy_f4(sDW, p.wdw[br], K * DWR, G * 128);
                publish_smem();
                group_bar(1, G * 128);
                loaded_branch = br;
            }
            const int q = grp * 128 + g.gtid;               // pair index inside the item
            const bool valid0 = q < PPI * nimg;
            const int im = valid0 ? q / PPI : 0;
            const int qi = valid0 ? q - im * PPI : 0;
            const int oy = qi / Wp, ox = 2 * (qi - oy * Wp);
            const bool valid1 = valid0 && ox + 1 < W;
            const int woff = im * PS + oy * WS + ox;        // even: the three LDS.64 of a window row are aligned
            RowSink<false, DENSE, true> sink0, sink1;
            sink0.scale = scale; sink0.shift = shift; sink0.valid = valid0;
            if (!DENSE) {
                sink0.obase = p.out[br].base + (long long)(n0 + im) * p.out[br].sN + p.out[br].org + oy * p.out[br].Ws + ox;
                sink0.tout = nullptr; sink0.sCo = (unsigned)p.out[br].sC; sink0.nout = p.nout;
            } else {
                sink0.dA = p.dstA[br]; sink0.dB = p.dstB[br]; sink0.split = p.split[br]; sink0.M = p.M[br];
                sink0.n = n0 + im; sink0.HW = HW; sink0.opix = oy * W + ox;
            }
            sink1 = sink0;
            sink1.valid = valid1;
            if (!DENSE) sink1.obase = sink0.obase + 1; else sink1.opix = sink0.opix + 1;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c, ++it) {
                const uint32_t buf = it % NB;
                mbar_wait(&fullb[buf], (it / NB) & 1u);
                float a0[8], a1[8];
                dw5_pair8<true>(X + (size_t)buf * BUF + woff, CS, WS, sDW + c * 8 * DWR, valid0, valid1, a0, a1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&freeb[buf]);
                put_chunk2<K, NP>(g, a0, a1, c, b_hi, b_lo);
            }
            get_tiles2<NP>(g, sink0, sink1, DENSE ? p.M[br] : p.nout);
        }
    }
    cta_teardown<TOT>(&tmem_slot);
}

// ===================================================================================================
// tc_head2w_kernel: tc_head2_kernel with SIBLING WARPS.  ncu on tc_head2_kernel (profiles/r2_kernels_ncu.txt): 9 warps per SM,
// issue slots 29 % busy, stalls spread over long/short scoreboard and MIO throttle: two warps per scheduler cannot hide the
// LDS -> FFMA chains of the 5x5 stencil, and TMEM (2 groups x 256 columns) rules out more warpgroups.  A warp may only touch
// the TMEM lanes 32 (warp % 4) .. +31, but nothing says ONE warp per lane quarter: here warps w and w + 4 of a group share the
// same 32 pixel pairs and each computes half of a chunk's channels (4 of 8: columns [4 sub, 4 sub + 4) of the hi / lo blocks),
// the chunk's MMAs are issued by the last of the group's EIGHT warps, and in the epilogue sibling 0 drains tile 0 (the pair's
// left pixel) while sibling 1 drains tile 1.  17 warps per SM on the same TMEM and shared-memory footprint.
// ===================================================================================================
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
template <bool RELU>
__device__ __forceinline__ void dw5_pair4(const float* __restrict__ xk, int RS, int WS, const float* __restrict__ wk, bool valid0, bool valid1,
                                          float (&a0)[4], float (&a1)[4]) {
    if (!valid0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
        return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float w[28];
#pragma unroll
        for (int t = 0; t < 7; ++t) {
            const float4 w4 = *reinterpret_cast<const float4*>(wk + 4 * t);
            w[4 * t] = w4.x; w[4 * t + 1] = w4.y; w[4 * t + 2] = w4.z; w[4 * t + 3] = w4.w;
        }
        float p0[5], p1[5];
        const float* row = xk;
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) {
            const float2 v01 = *reinterpret_cast<const float2*>(row);
            const float2 v23 = *reinterpret_cast<const float2*>(row + 2);
            const float2 v45 = *reinterpret_cast<const float2*>(row + 4);
            const float v[6] = {v01.x, v01.y, v23.x, v23.y, v45.x, v45.y};
            p0[dy] = w[dy * 5] * v[0];
            p1[dy] = w[dy * 5] * v[1];
#pragma unroll
            for (int dx = 1; dx < 5; ++dx) {
                p0[dy] = fmaf(w[dy * 5 + dx], v[dx], p0[dy]);
                p1[dy] = fmaf(w[dy * 5 + dx], v[dx + 1], p1[dy]);
            }
            row += WS;
        }
        float d0 = ((p0[0] + p0[1]) + (p0[2] + p0[3])) + p0[4];      // same association as dw8p
        float d1 = ((p1[0] + p1[1]) + (p1[2] + p1[3])) + p1[4];
        d0 = fmaf(d0, w[25], w[26]);
        d1 = fmaf(d1, w[25], w[26]);
        if (RELU) { d0 = fmaxf(d0, 0.f); d1 = fmaxf(d1, 0.f); }
        a0[j] = d0;
        a1[j] = valid1 ? d1 : 0.f;
        xk += RS;
        wk += 28;
    }
}
// four channels (sibling `sub`) of chunk c of both tiles -> TMEM; the last of the group's eight warps issues the chunk's MMAs
template <int KP, int NP>
__device__ __forceinline__ void put_chunk2w(Grp& g, int sub, const float (&a0)[4], const float (&a1)[4], int c, uint32_t b_hi, uint32_t b_lo) {
    const uint32_t buf = g.chunk & 1u, use = g.chunk >> 1;
    if (use > 0) mbar_wait(&g.pipe->empty[buf], (use - 1) & 1u);
    fence_after_sync();
    {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { hi[i] = __float_as_uint(a0[i]) & 0xFFFFE000u; lo[i] = __float_as_uint(a0[i] - __uint_as_float(hi[i])); }
        tmem_st4(g.tlane + buf * 32 + 4 * sub, hi);
        tmem_st4(g.tlane + buf * 32 + 8 + 4 * sub, lo);
#pragma unroll
        for (int i = 0; i < 4; ++i) { hi[i] = __float_as_uint(a1[i]) & 0xFFFFE000u; lo[i] = __float_as_uint(a1[i] - __uint_as_float(hi[i])); }
        tmem_st4(g.tlane + buf * 32 + 16 + 4 * sub, hi);
        tmem_st4(g.tlane + buf * 32 + 24 + 4 * sub, lo);
    }
    wait_st();
    fence_before_sync();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const uint32_t old = atom_add_acq_rel(&g.pipe->arrivals[buf], 1u);
        if ((old & 7u) == 7u) {
            fence_after_sync();
            constexpr uint32_t idesc = make_idesc_tf32(128, NP);
            constexpr uint32_t LBO = 128, SBO = (KP / 4) * 128;
            const uint64_t bh = make_b_desc(b_hi + c * 256, LBO, SBO);
            const uint64_t bl = make_b_desc(b_lo + c * 256, LBO, SBO);
#pragma unroll
            for (int tile = 0; tile < 2; ++tile) {
                const uint32_t a_hi = g.tcol + buf * 32 + tile * 16, a_lo = a_hi + 8, d = g.tcol + kA2Cols + tile * NP;
                mma_tf32_ts(d, a_lo, bh, idesc, c > 0 ? 1u : 0u);
                mma_tf32_ts(d, a_hi, bl, idesc, 1u);
                mma_tf32_ts(d, a_hi, bh, idesc, 1u);
            }
            mma_commit(&g.pipe->empty[buf]);
            if (c == KP / 8 - 1) mma_commit(&g.pipe->dfull);
        }
    }
    __syncwarp();
    ++g.chunk;
}

template <int K, int NP, bool DENSE>
__global__ void __launch_bounds__(2 * 256 + 32, 1)
tc_head2w_kernel(const __grid_constant__ HeadArgs p) {
    pdl_trigger();
    constexpr int G = 2, GT = 256, NCH = K / 8, NB = kHeadBufs, DWR = 28, COLS = 256, TOT = 512;
    static_assert(K % 8 == 0 && NP % 16 == 0 && kA2Cols + 2 * NP <= COLS, "shape");
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) Pipe pipes[G];
    __shared__ __align__(8) uint64_t fullb[NB], freeb[NB];
    __shared__ uint32_t tmem_slot;
    constexpr int WFL = 2 * NP * K + 2 * NP;
    float* sB = smem;
    float* sDW = sB + WFL;
    float* X = sDW + K * DWR;
    const int H = p.in[0].H, W = p.in[0].W, WS = p.in[0].Ws;
    // whole images: PS = one framed plane, `imgs` of them per item.  Banded (maps of more than 256 pixel pairs): an item is
    // band_rows rows of ONE image and PS = the band with its 2 + 2 halo rows (contiguous in the framed plane)
    const bool banded = p.band_rows > 0;
    const int PS = banded ? (p.band_rows + 4) * WS : (H + 4) * WS;
    const int CS = PS * p.imgs;
    const int BUF = 8 * CS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NB; ++i) { mbar_init(&fullb[i], 1); mbar_init(&freeb[i], G * 8); }
        for (int i = 0; i < G; ++i) {
            for (int b = 0; b < kMaxABufs; ++b) { mbar_init(&pipes[i].empty[b], 1); pipes[i].arrivals[b] = 0; }
            mbar_init(&pipes[i].dfull, 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, TOT);
    const int ngroups = banded ? p.N * p.bands : (p.N + p.imgs - 1) / p.imgs;
    const int items = 2 * ngroups;
    const int first_branch = (int)blockIdx.x / ngroups;
    if (threadIdx.x < G * GT) {
        copy_f4(sB, p.wpw[first_branch], WFL, G * GT);
        copy_f4(sDW, p.wdw[first_branch], K * DWR, G * GT);
        publish_smem();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    pdl_wait();
    uint32_t it = 0;
    if (warp == G * 8) {
        for (int t = blockIdx.x; t < items; t += gridDim.x) {
            const int br = t / ngroups, gi = t - br * ngroups;
            const int n0 = banded ? gi / p.bands : gi * p.imgs;
            const int nimg = banded ? 1 : min(p.imgs, p.N - n0);
            const int r0 = banded ? (gi % p.bands) * p.band_rows : 0;
            const int rows = banded ? min(p.band_rows, H - r0) : H;
            const uint32_t bytes = (uint32_t)((rows + 4) * WS * sizeof(float));      // (whole images: the whole framed plane)
            for (int c = 0; c < NCH; ++c, ++it) {
                const uint32_t buf = it % NB, use = it / NB;
                if (use > 0) mbar_wait(&freeb[buf], (use - 1) & 1u);
                publish_smem();
                if (lane == 0) mbar_expect_tx(&fullb[buf], 8u * nimg * bytes);
                __syncwarp();
                for (int j = lane; j < 8 * nimg; j += 32) {
                    const int ch = j & 7, i = j >> 3;
                    bulk_g2s(X + (size_t)buf * BUF + ch * CS + i * PS, plane_ptr(p.in[br], n0 + i, c * 8 + ch) + (size_t)r0 * WS, bytes, &fullb[buf]);
                }
            }
        }
    } else {
        const int grp = threadIdx.x >> 8, sub = (warp >> 2) & 1;
        Grp g;
        g.tcol = tmem_slot + grp * COLS;
        g.tlane = g.tcol + ((uint32_t)(32 * (warp & 3)) << 16);
        g.pipe = &pipes[grp];
        g.chunk = 0; g.dparity = 0;
        g.gtid = threadIdx.x & 127;
        const uint32_t b_hi = smem_u32(sB), b_lo = smem_u32(sB + NP * K);
        const float* scale = sB + 2 * NP * K;
        const float* shift = scale + NP;
        const int HW = H * W;
        const int Wp = (W + 1) >> 1;                        // pixel pairs per row
        int loaded_branch = first_branch;
        for (int t = blockIdx.x; t < items; t += gridDim.x) {
            const int br = t / ngroups, gi = t - br * ngroups;
            const int n0 = banded ? gi / p.bands : gi * p.imgs;
            const int nimg = banded ? 1 : min(p.imgs, p.N - n0);
            const int r0 = banded ? (gi % p.bands) * p.band_rows : 0;
            const int PPI = (banded ? min(p.band_rows, H - r0) : H) * Wp;       // pixel pairs per image of the item
            if (br != loaded_branch) {
                group_bar(1, G * GT);
                copy_f4(sB, p.wpw[br], WFL, G * GT);
                copy_f4(sDW, p.wdw[br], K * DWR, G * GT);
                publish_smem();
                group_bar(1, G * GT);
                loaded_branch = br;
            }
            const int q = grp * 128 + g.gtid;               // pair index inside the item
            bool valid0;
            int im, oyl, ox;                                // oyl: row inside the item's band
            if (p.use_map) {
                const unsigned int e = p.lanemap[q];
                im = (int)(e >> 16); oyl = (int)((e >> 8) & 0xFFu); ox = 2 * (int)(e & 0xFFu);
                valid0 = e != 0xFFFFFFFFu && im < nimg;
                if (!valid0) { im = 0; oyl = 0; ox = 0; }
            } else {
                valid0 = q < PPI * nimg;
                im = valid0 ? q / PPI : 0;
                const int qi = valid0 ? q - im * PPI : 0;
                oyl = qi / Wp; ox = 2 * (qi - oyl * Wp);
            }
            const int oy = r0 + oyl;
            const bool valid1 = valid0 && ox + 1 < W;
            const int woff = im * PS + oyl * WS + ox;       // even: the three LDS.64 of a window row are aligned
            // Epilogue: sibling `sub` drains ITS HALF of the accumulator columns of BOTH tiles, so the pair's two pixels leave as one 8-byte
            // store per channel (a sibling that drains one tile stores every other pixel: half-used sectors, twice the store instructions)
            float* const ob = DENSE ? nullptr
                                    : p.out[br].base + (long long)(n0 + im) * p.out[br].sN + p.out[br].org + oy * p.out[br].Ws + ox;
            const unsigned sCo = DENSE ? 0u : (unsigned)p.out[br].sC;
            const long long dn = n0 + im, dpix = oy * W + ox;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c, ++it) {
                const uint32_t buf = it % NB;
                mbar_wait(&fullb[buf], (it / NB) & 1u);
                float a0[4], a1[4];
                dw5_pair4<true>(X + (size_t)buf * BUF + 4 * sub * CS + woff, CS, WS, sDW + (c * 8 + 4 * sub) * DWR, valid0, valid1, a0, a1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&freeb[buf]);
                put_chunk2w<K, NP>(g, sub, a0, a1, c, b_hi, b_lo);
            }
            mbar_wait(&g.pipe->dfull, g.dparity);
            g.dparity ^= 1u;
            fence_after_sync();
            const int ncols = DENSE ? p.M[br] : p.nout;
            constexpr int H0 = (NP / 16 + 1) / 2 * 16;          // columns [0, H0) -> sibling 0, [H0, NP) -> sibling 1
            const int c_lo = sub ? H0 : 0, c_hi = min(sub ? NP : H0, ncols);
#pragma unroll 1
            for (int n0c = c_lo; n0c < c_hi; n0c += 16) {
                float d0[16], d1[16];
                tmem_ld16(g.tlane + kA2Cols + n0c, d0);
                tmem_ld16(g.tlane + kA2Cols + NP + n0c, d1);
                wait_ld();
                if (valid0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int m = n0c + j;
                        const float v0 = fmaf(d0[j], scale[m], shift[m]), v1 = fmaf(d1[j], scale[m], shift[m]);
                        float* dst = nullptr;
                        if (!DENSE) { if (m < p.nout) dst = ob + (unsigned)m * sCo; }
                        else if (m < p.split[br]) dst = p.dstA[br] + (dn * p.split[br] + m) * HW + dpix;
                        else if (m < p.M[br]) dst = p.dstB[br] + (dn * (p.M[br] - p.split[br]) + (m - p.split[br])) * HW + dpix;
                        if (dst) {
                            if (valid1 && p.st2_ok) *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);      // ox even: 8-byte aligned
                            else { *dst = v0; if (valid1) dst[1] = v1; }
                        }
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_slot, TOT);
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
__global__ void __launch_bounds__(G * 128, (G * (kACols + NP) > kTmemCols) ? 1 : 2)
tc_s1_kernel(const __grid_constant__ S1Args p) {
    pdl_trigger();
    constexpr int KP = K;
    constexpr int COLS = kACols + NP;
    constexpr int TOT = (G * COLS > kTmemCols) ? 512 : kTmemCols;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) Pipe pipes[G];
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
    publish_smem();
    Grp g = cta_setup<G, COLS, TOT>(pipes, &tmem_slot);
    pdl_wait();                                            // predecessor's activations are complete and visible from here on
    const uint32_t b1_hi = smem_u32(sB1), b1_lo = smem_u32(sB1 + NP * KP);
    const uint32_t b2_hi = smem_u32(sB2), b2_lo = smem_u32(sB2 + NP * KP);
    const float* sc1 = sB1 + 2 * NP * KP; const float* sh1 = sc1 + NP;
    const float* sc2 = sB2 + 2 * NP * KP; const float* sh2 = sc2 + NP;
    const int items = p.N * p.bandsPerImg;
    const int grp = threadIdx.x >> 7;
    const unsigned sC = (unsigned)p.P.sC;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int n = item / p.bandsPerImg;
        const int r0 = (item - n * p.bandsPerImg) * p.TR;
        const int rows = min(p.TR, H - r0);
        __syncthreads();
        // only the padding of T needs zeros: the two pad columns of every row and whole out-of-image halo rows
        for (int i = threadIdx.x; i < K * (rows + 2); i += G * 128) {
            const int k = i / (rows + 2), rr = i - k * (rows + 2);
            const int gr = r0 - 1 + rr;
            float* row = T + k * RS + rr * WS;
            if (gr < 0 || gr >= H) { for (int c = 0; c < WS; ++c) row[c] = 0.f; }
            else { row[0] = 0.f; row[W + 1] = 0.f; }
        }
        // ---- phase B: pw1 + BN + ReLU on every in-image pixel of rows [r0-1, r0+rows] -> T --------------------------
        const int gr_lo = max(r0 - 1, 0), gr_hi = min(r0 + rows, H - 1);
        const int npos = (gr_hi - gr_lo + 1) * W;
        for (int tile = grp; tile * 128 < npos; tile += G) {
            const int q = tile * 128 + g.gtid;
            const bool valid = q < npos;
            const int rr = valid ? q / W : 0, x = valid ? q - rr * W : 0;
            const int gr = gr_lo + rr;
            const float* ibase = p.P.base + (long long)n * p.P.sN + p.P.org + gr * p.P.Ws + x;
            float* tpos = T + (gr - (r0 - 1)) * WS + 1 + x;
            float v[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) v[k] = valid ? __ldg(ibase + p.tin.c[k] * sC) : 0.f;    // all loads in flight at once
            pw_tile<KP, NP>(g, b1_hi, b1_lo,
                [&](int k0, float (&a)[8]) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = v[k0 + j];
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
            float* obase = p.P.base + (long long)n * p.P.sN + p.P.org + (r0 + orow) * p.P.Ws + ox;
            pw_tile_rolled<KP, NP>(g, b2_hi, b2_lo,
                [&](int k0, float (&a)[8]) { dw8<3, 1, false>(win, RS, WS, sDW, k0, valid, a); },
                [&](int n0, float (&d)[16]) {
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + j < K) obase[p.tout.c[n0 + j] * sC] = fmaxf(fmaf(d[j], sc2[n0 + j], sh2[n0 + j]), 0.f);
                    }
                });
        }
    }
    cta_teardown<TOT>(&tmem_slot);
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
__global__ void __launch_bounds__(G * 128, (G * (kACols + NP) > kTmemCols) ? 1 : 2)
tc_s2_kernel(const __grid_constant__ S2Args p) {
    pdl_trigger();
    constexpr int KP = K;
    constexpr int COLS = kACols + NP;
    constexpr int TOT = (G * COLS > kTmemCols) ? 512 : kTmemCols;
    static_assert(K <= G * 128, "one bulk copy per thread");
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) Pipe pipes[G];
    __shared__ __align__(8) uint64_t xbar;
    __shared__ uint32_t tmem_slot;
    constexpr int WFL = 2 * NP * KP + 2 * NP;
    float* sBp = smem;
    float* sB1 = sBp + WFL;
    float* sB2 = sB1 + WFL;
    float* sDWp = sB2 + WFL;
    float* sDWm = sDWp + K * 12;
    float* X = sDWm + K * 12;
    const int Hin = p.in.H, Win = p.in.W, WS = p.in.Ws;
    const int Hout = p.out.H, Wout = p.out.W;
    const int RS = (2 * p.TR + 1) * WS;
    const int pin = p.in.pad;                 // frame pad of the input pool (>= 1)
    copy_f4(sBp, p.wp, WFL, G * 128);
    copy_f4(sB1, p.w1, WFL, G * 128);
    copy_f4(sB2, p.w2, WFL, G * 128);
    copy_f4(sDWp, p.wdwp, K * 12, G * 128);
    copy_f4(sDWm, p.wdwm, K * 12, G * 128);
    publish_smem();
    if (threadIdx.x == 0) { mbar_init(&xbar, 1); fence_mbar_init(); }
    Grp g = cta_setup<G, COLS, TOT>(pipes, &tmem_slot);
    pdl_wait();                                            // predecessor's activations are complete and visible from here on
    const uint32_t bp_hi = smem_u32(sBp), bp_lo = smem_u32(sBp + NP * KP);
    const uint32_t b1_hi = smem_u32(sB1), b1_lo = smem_u32(sB1 + NP * KP);
    const uint32_t b2_hi = smem_u32(sB2), b2_lo = smem_u32(sB2 + NP * KP);
    const float* scp = sBp + 2 * NP * KP; const float* shp = scp + NP;
    const float* sc1 = sB1 + 2 * NP * KP; const float* sh1 = sc1 + NP;
    const float* sc2 = sB2 + 2 * NP * KP; const float* sh2 = sc2 + NP;
    const int items = p.N * p.bandsPerImg;
    const int grp = threadIdx.x >> 7;
    const unsigned sCo = (unsigned)p.out.sC;
    uint32_t xparity = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int n = item / p.bandsPerImg;
        const int r0 = (item - n * p.bandsPerImg) * p.TR;
        const int rows = min(p.TR, Hout - r0);
        const int gr0 = 2 * r0 - 1, nrows = 2 * rows + 1;       // image rows [gr0, gr0+nrows) <-> staged rows [0, nrows)
        __syncthreads();
        publish_smem();
        if (threadIdx.x == 0) mbar_expect_tx(&xbar, (uint32_t)(K * nrows * WS * sizeof(float)));
        stage_bulk<K>(X, RS, p.in, p.tin, n, gr0 + pin, nrows, &xbar, 0);
        mbar_wait(&xbar, xparity);
        xparity ^= 1u;
        const int npix = rows * Wout;
        const int coff = pin - 1;
        // ---- proj: dw3x3 s2 + BN -> pw + BN + ReLU on the raw input ---------------------------------------------
        for (int tile = grp; tile * 128 < npix; tile += G) {
            const int q = tile * 128 + g.gtid;
            const bool valid = q < npix;
            const int orow = valid ? q / Wout : 0, ox = valid ? q - orow * Wout : 0;
            const float* win = X + (2 * orow) * WS + 2 * ox + coff;
            float* obase = p.out.base + (long long)n * p.out.sN + p.out.org + (r0 + orow) * p.out.Ws + ox;
            pw_tile_rolled<KP, NP>(g, bp_hi, bp_lo,
                [&](int k0, float (&a)[8]) { dw8<3, 2, false>(win, RS, WS, sDWp, k0, valid, a); },
                [&](int n0, float (&d)[16]) {
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + j < K) obase[p.tout.c[n0 + j] * sCo] = fmaxf(fmaf(d[j], scp[n0 + j], shp[n0 + j]), 0.f);
                    }
                });
        }
        __syncthreads();
        // ---- main pw1 in place on every staged in-image pixel (frame / halo positions stay zero) ------------------------
        const int gr_lo = max(gr0, 0), gr_hi = min(gr0 + nrows - 1, Hin - 1);
        const int npos = (gr_hi - gr_lo + 1) * Win;
        for (int tile = grp; tile * 128 < npos; tile += G) {
            const int q = tile * 128 + g.gtid;
            const bool valid = q < npos;
            const int rr = valid ? q / Win : 0, x = valid ? q - rr * Win : 0;
            float* tpos = X + (gr_lo + rr - gr0) * WS + pin + x;
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
            const float* win = X + (2 * orow) * WS + 2 * ox + coff;
            float* obase = p.out.base + (long long)n * p.out.sN + p.out.org + (r0 + orow) * p.out.Ws + ox;
            pw_tile_rolled<KP, NP>(g, b2_hi, b2_lo,
                [&](int k0, float (&a)[8]) { dw8<3, 2, false>(win, RS, WS, sDWm, k0, valid, a); },
                [&](int n0, float (&d)[16]) {
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + j < K) obase[p.tout.c[K + n0 + j] * sCo] = fmaxf(fmaf(d[j], sc2[n0 + j], sh2[n0 + j]), 0.f);
                    }
                });
        }
    }
    cta_teardown<TOT>(&tmem_slot);
}

// ===================================================================================================
// tc_stem_kernel: conv3x3 s2 p1 (3->24) + BN + ReLU + maxpool3x3 s2 p1 (reference shufflenetv2.py:74-80,103-104) with the
// convolution as an implicit GEMM on the tensor cores: M = conv positions (thread = position), K = 27 taps (padded to 32),
// N = 24 channels (padded to 32).  A work item is (image, band of TRo pooled rows, tile of TWo pooled columns): the input
// patch is staged by one TMA bulk copy per (channel, row) [uint8 input: converted with the reference's /255 on the way in],
// the conv tile goes through shared memory once, the 3x3/s2 max-pool reads it back and writes the framed output planes.
// ===================================================================================================
struct StemTcArgs {
    const void* x;          // [N,3,H,W] fp32 or uint8
    Planes out;             // 24 planes at H/4 x W/4
    const float* wpack;     // tc pack K=27->32, N=24->32
    int N, H, W;
    int TRo, TWo, tilesX, tilesY;
};

template <bool U8>
__global__ void __launch_bounds__(512, 2)
tc_stem_kernel(const __grid_constant__ StemTcArgs p) {
    pdl_trigger();
    constexpr int G = 4, KP = 32, NP = 32, COLS = kACols + NP;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) Pipe pipes[G];
    __shared__ __align__(8) uint64_t xbar[2];
    __shared__ uint32_t tmem_slot;
    const int TRo = p.TRo, TWo = p.TWo;
    const int IR = 4 * TRo + 3, Wst = 4 * TWo + 8;           // staged input rows / columns (column 0 <-> input column 4*ox0 - 4)
    const int CR = 2 * TRo + 1, CC = 2 * TWo + 1;            // conv tile
    constexpr int WFL = 2 * NP * KP + 2 * NP;
    float* sB = smem;
    float* XinBuf = sB + WFL;                                // 2 x [3][IR][Wst]: the next item's patch is staged while this one computes
    const int xsz = 3 * IR * Wst;
    float* Cv = XinBuf + 2 * xsz;                            // [24][CR*CC]
    copy_f4(sB, p.wpack, WFL, 512);
    publish_smem();
    if (threadIdx.x == 0) { mbar_init(&xbar[0], 1); mbar_init(&xbar[1], 1); fence_mbar_init(); }
    Grp g = cta_setup<G, COLS>(pipes, &tmem_slot);
    pdl_wait();                                            // predecessor's activations are complete and visible from here on
    const uint32_t b_hi = smem_u32(sB), b_lo = smem_u32(sB + NP * KP);
    const float* scale = sB + 2 * NP * KP;
    const float* shift = scale + NP;
    const int H = p.H, W = p.W, HC = H / 2, WC = W / 2, HO = H / 4, WO = W / 4;
    const int items = p.N * p.tilesX * p.tilesY;
    const int grp = threadIdx.x >> 7;

    // stage the input patch of `item` into buffer `b` (bulk copies for fp32, converting loads for uint8)
    auto stage = [&](int item, int b) {
        const int n = item / (p.tilesX * p.tilesY);
        const int rem = item - n * (p.tilesX * p.tilesY);
        const int ty = rem / p.tilesX, tx = rem - ty * p.tilesX;
        const int iy0 = 4 * (ty * TRo) - 3, ic0 = 4 * (tx * TWo) - 4;
        const int c_lo = max(ic0, 0), c_hi = min(ic0 + Wst, W);
        float* Xin = XinBuf + b * xsz;
        if (!U8) {
            int nvalid = 0;
            for (int r = 0; r < IR; ++r) nvalid += (iy0 + r >= 0 && iy0 + r < H);
            if (threadIdx.x == 0) mbar_expect_tx(&xbar[b], (uint32_t)(3 * nvalid * (c_hi - c_lo) * sizeof(float)));
            for (int i = threadIdx.x; i < 3 * IR; i += 512) {
                const int c = i / IR, r = i - c * IR;
                const int iy = iy0 + r;
                float* dst = Xin + (c * IR + r) * Wst;
                if (iy >= 0 && iy < H) {
                    bulk_g2s(dst + (c_lo - ic0), reinterpret_cast<const float*>(p.x) + (((size_t)n * 3 + c) * H + iy) * W + c_lo,
                             (uint32_t)((c_hi - c_lo) * sizeof(float)), &xbar[b]);
                    for (int j = 0; j < c_lo - ic0; ++j) dst[j] = 0.f;
                    for (int j = c_hi - ic0; j < Wst; ++j) dst[j] = 0.f;
                } else {
                    for (int j = 0; j < Wst; ++j) dst[j] = 0.f;
                }
            }
        } else {
            const int q4 = Wst / 4;
            for (int i = threadIdx.x; i < 3 * IR * q4; i += 512) {
                const int cr = i / q4, j4 = i - cr * q4;
                const int c = cr / IR, r = cr - c * IR;
                const int iy = iy0 + r, ix = ic0 + 4 * j4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
                    const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(reinterpret_cast<const uint8_t*>(p.x) + (((size_t)n * 3 + c) * H + iy) * W + ix));
                    v = make_float4(__fdiv_rn((float)u.x, 255.0f), __fdiv_rn((float)u.y, 255.0f), __fdiv_rn((float)u.z, 255.0f), __fdiv_rn((float)u.w, 255.0f));
                }
                *reinterpret_cast<float4*>(Xin + cr * Wst + 4 * j4) = v;
            }
        }
    };

    uint32_t xpar[2] = {0u, 0u};
    int it = 0;
    if (blockIdx.x < items) stage(blockIdx.x, 0);
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int cur = it & 1;
        const int n = item / (p.tilesX * p.tilesY);
        const int rem = item - n * (p.tilesX * p.tilesY);
        const int ty = rem / p.tilesX, tx = rem - ty * p.tilesX;
        const int oy0 = ty * TRo, ox0 = tx * TWo;
        const float* Xin = XinBuf + cur * xsz;
        // the other buffer was last read by the conv phase of the previous item, which every thread left before the
        // pooling barrier of that item: it is free, prefetch the next item into it
        if (item + gridDim.x < items) { publish_smem(); stage(item + gridDim.x, cur ^ 1); }
        if (!U8) { mbar_wait(&xbar[cur], xpar[cur]); xpar[cur] ^= 1u; }
        __syncthreads();                                      // edge zero-fill / uint8 conversion stores of this buffer are visible
        // ---- conv as GEMM: thread = conv position -----------------------------------------------------------------
        const int npos = CR * CC;
        for (int tile = grp; tile * 128 < npos; tile += G) {
            const int q = tile * 128 + g.gtid;
            const bool inb = q < npos;
            const int lr = inb ? q / CC : 0, lc = inb ? q - lr * CC : 0;
            const int cy = 2 * oy0 - 1 + lr, cx = 2 * ox0 - 1 + lc;
            const bool valid = inb && cy >= 0 && cy < HC && cx >= 0 && cx < WC;
            const float* xp = Xin + (2 * lr) * Wst + 2 * lc + 1;
            float v[KP];
            {
                const float* rp = xp;
                const int cstep = (IR - 3) * Wst;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) v[c * 9 + ky * 3 + kx] = valid ? rp[kx] : 0.f;
                        rp += Wst;
                    }
                    rp += cstep;
                }
            }
#pragma unroll
            for (int k = 27; k < KP; ++k) v[k] = 0.f;
            pw_tile<KP, NP>(g, b_hi, b_lo,
                [&](int k0, float (&a)[8]) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = v[k0 + j];
                },
                [&](int n0, float (&d)[16]) {
                    if (inb) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + j < 24) Cv[(n0 + j) * npos + q] = valid ? fmaxf(fmaf(d[j], scale[n0 + j], shift[n0 + j]), 0.f) : -INFINITY;
                    }
                });
        }
        __syncthreads();
        // ---- maxpool 3x3 s2 p1 (PyTorch pads with -inf) -> framed output planes ----------------------------------------
        const int rows = min(TRo, HO - oy0), cols = min(TWo, WO - ox0);
        {   // warp <-> channel, lane <-> pooled pixel: no integer division in the hot loop
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            for (int ch = warp; ch < 24; ch += 16) {
                float* op = plane_ptr(p.out, n, ch) + p.out.org + oy0 * p.out.Ws + ox0;
                const float* cbase = Cv + ch * npos;
                for (int oyl = 0; oyl < rows; ++oyl)
                    for (int oxl = lane; oxl < cols; oxl += 32) {
                        const float* cv = cbase + (2 * oyl) * CC + 2 * oxl;
                        float m = fmaxf(fmaxf(cv[0], cv[1]), cv[2]);
                        m = fmaxf(m, fmaxf(fmaxf(cv[CC], cv[CC + 1]), cv[CC + 2]));
                        m = fmaxf(m, fmaxf(fmaxf(cv[2 * CC], cv[2 * CC + 1]), cv[2 * CC + 2]));
                        op[oyl * p.out.Ws + oxl] = m;
                    }
            }
        }
        __syncthreads();                                      // Cv is rewritten by the next item's conv phase
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

int tc_launch_stem(const void* x, int is_u8, const Planes& out, const float* wpack, int N, int H, int W, cudaStream_t s) {
    StemTcArgs a{x, out, wpack, N, H, W, 0, 0, 0, 0};
    const int HO = H / 4, WO = W / 4;
    a.TRo = HO >= 2 ? 2 : 1;
    a.TWo = WO < 44 ? WO : 44;
    while (a.TWo > 4 && (size_t)(2 * 32 * 32 + 64 + 2 * 3 * (4 * a.TRo + 3) * (4 * a.TWo + 8) + 24 * (2 * a.TRo + 1) * (2 * a.TWo + 1)) * sizeof(float) > 110 * 1024) --a.TWo;
    a.tilesX = (WO + a.TWo - 1) / a.TWo; a.tilesY = (HO + a.TRo - 1) / a.TRo;
    const size_t bytes = (size_t)(2 * 32 * 32 + 64 + 2 * 3 * (4 * a.TRo + 3) * (4 * a.TWo + 8) + 24 * (2 * a.TRo + 1) * (2 * a.TWo + 1) + 4) * sizeof(float);
    const int items = N * a.tilesX * a.tilesY;
    if (is_u8) {
        TRYL(set_smem_attr(tc_stem_kernel<true>, bytes));
        YFV2_CUDA(launch_k(tc_stem_kernel<true>, min(items, 2 * sm_count()), 512, bytes, s, pdl_take(), a));
    } else {
        TRYL(set_smem_attr(tc_stem_kernel<false>, bytes));
        YFV2_CUDA(launch_k(tc_stem_kernel<false>, min(items, 2 * sm_count()), 512, bytes, s, pdl_take(), a));
    }
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}

int tc_launch_s1(int K, const Planes& P, const ChanTab& tin, const ChanTab& tout, const float* w1, const float* wdw, const float* w2,
                 int N, cudaStream_t s) {
    S1Args a{P, tin, tout, w1, wdw, w2, N, 0, 0};
    const int H = P.H, W = P.W;
    auto run = [&](auto kern, int KK, int NP, int G, bool one_per_sm) -> int {
        const size_t wfl = (size_t)2 * (2 * NP * KK + 2 * NP) + KK * 12;
        auto bytes = [&](int tr) { return (wfl + (size_t)KK * (tr + 2) * (W + 2) + 4) * sizeof(float); };
        const size_t cap = one_per_sm ? 215 * 1024 : 110 * 1024;       // one or two CTAs per SM
        int TR = H;
        while (TR > 1 && bytes(TR) > cap) TR = (TR + 1) / 2;
        a.TR = TR; a.bandsPerImg = (H + TR - 1) / TR;
        TRYL(set_smem_attr(kern, bytes(TR)));
        const int items = N * a.bandsPerImg;
        YFV2_CUDA(launch_k(kern, min(items, (one_per_sm ? 1 : 2) * sm_count()), G * 128, bytes(TR), s, pdl_take(), a));
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    // (measured: whole-image items with 4 groups on one CTA per SM are 5 % slower for K=48 than two half-image CTAs)
    if (K == 24) return run(tc_s1_kernel<24, 32, 4>, 24, 32, 4, false);
    if (K == 48) return run(tc_s1_kernel<48, 48, 3>, 48, 48, 3, false);
    set_error("tc_launch_s1: unsupported K=%d", K);
    return YFV2_EUNSUPPORTED;
}

int tc_launch_s2(int K, const Planes& in, const Planes& out, const ChanTab& tin, const ChanTab& tout, const float* wdwp, const float* wp,
                 const float* w1, const float* wdwm, const float* w2, int N, cudaStream_t s) {
    S2Args a{in, out, tin, tout, wdwp, wp, w1, wdwm, w2, N, 0, 0};
    const int Hout = out.H;
    auto run = [&](auto kern, int KK, int NP, int G, bool one_per_sm) -> int {
        const size_t wfl = (size_t)3 * (2 * NP * KK + 2 * NP) + 2 * KK * 12;
        auto bytes = [&](int tr) { return (wfl + (size_t)KK * (2 * tr + 1) * in.Ws + 4) * sizeof(float); };
        const size_t cap = one_per_sm ? 218 * 1024 : 110 * 1024;
        int TR = Hout;
        while (TR > 1 && bytes(TR) > cap) --TR;
        a.TR = TR; a.bandsPerImg = (Hout + TR - 1) / TR;
        TRYL(set_smem_attr(kern, bytes(TR)));
        const int items = N * a.bandsPerImg;
        YFV2_CUDA(launch_k(kern, min(items, (one_per_sm ? 1 : 2) * sm_count()), G * 128, bytes(TR), s, pdl_take(), a));
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    // K=48: three 18.8 KB weight packs leave two co-resident CTAs only 2-row bands (one third-full tile per phase);
    // one CTA per SM with 8-row bands and 4 groups is 15 % faster (201 -> 171 us).  K=24: the opposite (197 vs 229 us).
    if (K == 24) return run(tc_s2_kernel<24, 32, 4>, 24, 32, 4, false);
    if (K == 48) return run(tc_s2_kernel<48, 48, 4>, 48, 48, 4, true);
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
        const int per_sm = bytes <= 110 * 1024 ? 2 : 1;
        YFV2_CUDA(launch_k(kern, min((ntiles + G - 1) / G, per_sm * sm_count()), G * 128, bytes, s, pdl_take(), a));
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    if (kind == 0) return run(tc_pw_kernel<96, 0, 0, 96, 2, true>, 96, 96, 2, 96);
    if (kind == 1) return run(tc_pw_kernel<192, 0, 0, 80, 2, true>, 192, 80, 2, 72);
    static const bool nb3 = getenv("YFV2_FPN_NB3") != nullptr;    // experiment: 3-deep A ring (48 + 80 = 128 columns per group)
    if (kind == 2 && nb3) return run(tc_pw_kernel<192, 96, 1, 80, 4, true, 3>, 288, 80, 4, 72);
    if (kind == 2) return run(tc_pw_kernel<192, 96, 1, 80, 4, true>, 288, 80, 4, 72);   // 184 KB of weights: one CTA per SM, so 4 groups
    set_error("tc_launch_pw: unknown kind %d", kind);
    return YFV2_EUNSUPPORTED;
}

// Band height / images per item for the DW->PW kernels: whole images (several per item) when one fits comfortably,
// otherwise row bands of one image.
static void dwpw_geometry(DwPwArgs& a, int Hout, size_t wfl_floats, size_t plane_floats_per_row, int halo_rows, int S, int Wout, int G) {
    const size_t cap = 200 * 1024 / sizeof(float);
    auto band_floats = [&](int tr) { return plane_floats_per_row * (size_t)(S * (tr - 1) + halo_rows); };
    int TR = Hout;
    while (TR > 1 && wfl_floats + band_floats(TR) > cap) --TR;
    a.TR = TR; a.bandsPerImg = (Hout + TR - 1) / TR; a.imgs = 1;
    if (a.bandsPerImg == 1) {
        const int tiles1 = (Hout * Wout + 127) / 128;                 // tiles one image needs
        while (a.imgs * tiles1 < G && wfl_floats + band_floats(TR) * (a.imgs + 1) <= cap) ++a.imgs;
    }
}

// K=96 DW3x3(stride)->PW (+ReLU) for nbranch branches (stage-4 blocks)
// K = 96 stride-2 depthwise -> pointwise branches on output maps of at most 512 pixels: channel-streamed whole-image items.
// Returns YFV2_EUNSUPPORTED (without setting an error the caller reports) when the geometry does not fit; the caller falls back.
bool tc_dws2c_supported(int K, const Planes& in, const Planes& out, int N, int* imgs_out, int* G_out, size_t* bytes_out) {
    const int HWo = out.H * out.W;
    if (HWo > 4 * 128 || in.pad < 1 || (K != 96 && K != 48)) return false;
    const size_t PS = (size_t)(in.H + 2 * in.pad) * in.Ws;
    const size_t wfl = (size_t)(2 * K * K + 2 * K) + K * 12;
    int imgs = 1;
    while ((imgs + 1) * HWo <= 4 * 128 && imgs + 1 <= N && (wfl + kDwsBufs * 4 * PS * (imgs + 1) + 4) * sizeof(float) <= kSmemCap - 2048) ++imgs;
    const size_t bytes = (wfl + kDwsBufs * 4 * PS * imgs + 4) * sizeof(float);
    if (bytes > kSmemCap - 2048) return false;
    *imgs_out = imgs; *G_out = (imgs * HWo + 127) / 128; *bytes_out = bytes;
    return true;
}
int tc_launch_dws2c(int K, int nbranch, const Planes* in, const ChanTab* tin, const Planes* out, const ChanTab* tout,
                    const float* const* wdw, const float* const* wpw, int N, cudaStream_t s) {
    Dws2Args a{};
    for (int b = 0; b < nbranch; ++b) { a.in[b] = in[b]; a.out[b] = out[b]; a.tin[b] = tin[b]; a.tout[b] = tout[b]; a.wdw[b] = wdw[b]; a.wpw[b] = wpw[b]; }
    a.N = N; a.nbranch = nbranch; a.nout = K;
    int G = 0; size_t bytes = 0;
    if (!tc_dws2c_supported(K, in[0], out[0], N, &a.imgs, &G, &bytes)) { set_error("tc_launch_dws2c: unsupported geometry"); return YFV2_EUNSUPPORTED; }
    const int items = nbranch * ((N + a.imgs - 1) / a.imgs);
    static const bool scalar_rows = getenv("YFV2_DWS2_SCALAR") != nullptr;          // A/B: three scalar loads per window row
    a.pair_rows = (!scalar_rows && (in[0].pad & 1) && (nbranch < 2 || (in[1].pad & 1))) ? 1 : 0;
    auto run = [&](auto kern, int g) -> int {
        TRYL(set_smem_attr(kern, bytes));
        YFV2_CUDA(launch_k(kern, min(items, sm_count()), g * 128 + 32, bytes, s, pdl_take(), a));
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    if (K == 48) {
        if (G <= 1) return run(tc_dws2c_kernel<48, 48, 1>, 1);
        if (G == 2) return run(tc_dws2c_kernel<48, 48, 2>, 2);
        if (G == 3) return run(tc_dws2c_kernel<48, 48, 3>, 3);
        return run(tc_dws2c_kernel<48, 48, 4>, 4);
    }
    if (G <= 1) return run(tc_dws2c_kernel<96, 96, 1>, 1);
    if (G == 2) return run(tc_dws2c_kernel<96, 96, 2>, 2);
    if (G == 3) return run(tc_dws2c_kernel<96, 96, 3>, 3);
    return run(tc_dws2c_kernel<96, 96, 4>, 4);
}

int tc_launch_dwpw96(int stride, int nbranch, const Planes* in, const ChanTab* tin, const Planes* out, const ChanTab* tout,
                     const float* const* wdw, const float* const* wpw, int N, cudaStream_t s) {
    DwPwArgs a{};
    for (int b = 0; b < nbranch; ++b) { a.in[b] = in[b]; a.out[b] = out[b]; a.tin[b] = tin[b]; a.tout[b] = tout[b]; a.wdw[b] = wdw[b]; a.wpw[b] = wpw[b]; }
    a.N = N; a.nbranch = nbranch; a.nout = 96;
    const int Hout = out[0].H;
    constexpr int G = 2;
    auto run = [&](auto kern, int S) -> int {
        const size_t wfl = (size_t)(2 * 96 * 96 + 2 * 96) + 96 * 12;
        dwpw_geometry(a, Hout, wfl + 4, (size_t)96 * in[0].Ws, 3, S, out[0].W, G);
        const size_t bytes = (wfl + (size_t)96 * (S * (a.TR - 1) + 3) * in[0].Ws * a.imgs + 4) * sizeof(float);
        TRYL(set_smem_attr(kern, bytes));
        const int items = ((N + a.imgs - 1) / a.imgs) * a.bandsPerImg * nbranch;
        YFV2_CUDA(launch_k(kern, min(items, sm_count()), G * 128, bytes, s, pdl_take(), a));
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    if (stride == 1) return run(tc_dwpw_kernel<96, 96, G, 3, 1, false, true, false>, 1);
    return run(tc_dwpw_kernel<96, 96, G, 3, 2, false, true, false>, 2);
}

// heads: half 0: T = BN(pw(ReLU(BN(dw5x5(S)))));  half 1: preds = F(ReLU(BN(dw5x5(T)))) with F = outconv o BN o pw folded
// into one [M x 72] matrix + bias at pack time (plan.cu).  branch 0 = cls head (outputs obj+cls), branch 1 = reg head.
// Deal the H x ceil(W/2) pixel pairs of `imgs` framed planes (row stride WS, plane stride PS floats) to 256 lanes so that the 16
// lanes of every half-warp read 16 different 8-byte shared-memory banks: bank unit of a pair = (offset / 2) mod 16, the window
// taps shift all lanes alike.  Greedy in raster order (keeps a half-warp's pairs close together for the epilogue's stores);
// pairs that cannot be placed conflict-free (a residue class with more than 16 members) fill the remaining lanes.
static void build_lanemap(HeadArgs& a, int H, int W, int WS, int imgs, size_t PS) {
    const int Wp = (W + 1) / 2, total = imgs * H * Wp;
    static thread_local int key[5] = {0, 0, 0, 0, 0};
    static thread_local unsigned int cached[256];
    if (key[0] != H || key[1] != W || key[2] != WS || key[3] != imgs || key[4] != (int)PS) {
        std::vector<unsigned int> pend((size_t)total), rest;
        for (int i = 0; i < total; ++i) {
            const int im = i / (H * Wp), r = (i / Wp) % H, j = i % Wp;
            pend[(size_t)i] = ((unsigned)im << 16) | ((unsigned)r << 8) | (unsigned)j;
        }
        auto unit = [&](unsigned int e) { return (int)((((size_t)(e >> 16) * PS + (size_t)((e >> 8) & 0xFFu) * WS + 2 * (size_t)(e & 0xFFu)) / 2) % 16); };
        for (int i = 0; i < 256; ++i) cached[i] = 0xFFFFFFFFu;
        for (int h = 0; h < 16; ++h) {
            bool used[16] = {false};
            int got = 0;
            rest.clear();
            for (unsigned int e : pend) {
                const int u = unit(e);
                if (got < 16 && !used[u]) { used[u] = true; cached[16 * h + got++] = e; }
                else rest.push_back(e);
            }
            pend.swap(rest);
        }
        for (int i = 0; i < 256 && !pend.empty(); ++i)
            if (cached[i] == 0xFFFFFFFFu) { cached[i] = pend.back(); pend.pop_back(); }
        key[0] = H; key[1] = W; key[2] = WS; key[3] = imgs; key[4] = (int)PS;
    }
    memcpy(a.lanemap, cached, sizeof(cached));
    a.use_map = 1;
}

static int heads_st2_ok(const HeadArgs& a, int half, int W) {
    if (W & 1) return 0;
    auto al8 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 7u) == 0; };
    if (half == 0) {
        for (int b = 0; b < 2; ++b)
            if (!al8(a.out[b].base) || (a.out[b].sN & 1) || (a.out[b].sC & 1) || (a.out[b].org & 1) || (a.out[b].Ws & 1)) return 0;
        return 1;
    }
    for (int b = 0; b < 2; ++b) if (!al8(a.dstA[b]) || !al8(a.dstB[b])) return 0;
    return 1;
}

int tc_launch_heads(int half, const Planes& sIn, const Planes& tcls, const Planes& treg, const float* const wdw[2], const float* const wpw[2],
                    float* reg, float* obj, float* cls, int A, int C, int N, cudaStream_t s) {
    if (A + C > 96 || 4 * A > 96) { set_error("tc heads: A+C=%d exceeds the output tile (96)", A + C); return YFV2_EUNSUPPORTED; }
    if (sIn.pad != 2) { set_error("tc heads: the 5x5 stencil needs input planes framed by 2"); return YFV2_EINVAL; }
    const int H = sIn.H, W = sIn.W;
    const int np = half == 0 ? 80 : 96;
    static const bool force_band = getenv("YFV2_HEADS_BAND") != nullptr;
    // ---- fast path: channel-streamed whole-image items ----------------------------------------------------------
    static const bool force_g4 = getenv("YFV2_HEADS_G4") != nullptr;
    // pairs pay off when every pair is full (even W); odd maps (11x11) keep one pixel per thread
    // odd maps (11x11 at 352x352) keep one pixel per thread: the pair kernel works on them (YFV2_HEADS_ODD_PAIRS=1, parity green) but
    // its half-empty last pairs cost what the shared window rows save (measured: heads3.a 62.8 vs 59.5 us, heads3.b 60.9 vs 62.0 us)
    static const bool odd_pairs = getenv("YFV2_HEADS_ODD_PAIRS") != nullptr;
    if ((W % 2 == 0 || odd_pairs) && H * ((W + 1) / 2) <= 256 && !force_band && !force_g4) {
        // pixel pairs: 2 warpgroups x 2 tiles
        HeadArgs a{};
        a.N = N; a.nout = 72;
        for (int b = 0; b < 2; ++b) { a.wdw[b] = wdw[b]; a.wpw[b] = wpw[b]; }
        if (half == 0) { a.in[0] = sIn; a.in[1] = sIn; a.out[0] = tcls; a.out[1] = treg; }
        else {
            a.in[0] = tcls; a.in[1] = treg;
            a.dstA[0] = obj; a.dstB[0] = cls; a.split[0] = A; a.M[0] = A + C;
            a.dstA[1] = reg; a.dstB[1] = reg; a.split[1] = 4 * A; a.M[1] = 4 * A;
        }
        const size_t PS = (size_t)(H + 4) * sIn.Ws;
        const size_t wfl = (size_t)(2 * np * 72 + 2 * np) + 72 * 28;
        const int PPI = H * ((W + 1) / 2);
        a.st2_ok = heads_st2_ok(a, half, W);
        a.imgs = 1;
        while ((a.imgs + 1) * PPI <= 256 && a.imgs + 1 <= N &&
               (wfl + kHeadBufs * 8 * PS * (a.imgs + 1) + 4) * sizeof(float) <= kSmemCap - 1024) ++a.imgs;
        const size_t bytes = (wfl + kHeadBufs * 8 * PS * a.imgs + 4) * sizeof(float);
        static const bool raster_heads = getenv("YFV2_HEADS_RASTER") != nullptr;      // A/B: raster lane order (bank conflicted)
        if (!raster_heads && H < 256 && W / 2 < 256 && a.imgs < 0xFFFF) build_lanemap(a, H, W, sIn.Ws, a.imgs, PS);
        if (bytes <= kSmemCap - 1024) {
            const int ngroups = (N + a.imgs - 1) / a.imgs;
            static const bool old_heads = getenv("YFV2_HEADS_OLD") != nullptr;     // round-1 kernel (one warp per lane quarter), kept for A/B runs
            auto run = [&](auto kern, int threads) -> int {
                TRYL(set_smem_attr(kern, bytes));
                YFV2_CUDA(launch_k(kern, min(2 * ngroups, sm_count()), threads, bytes, s, pdl_take(), a));
                YFV2_LAUNCH_CHECK();
                return YFV2_OK;
            };
            if (old_heads) {
                if (half == 0) return run(tc_head2_kernel<72, 80, false>, 2 * 128 + 32);
                return run(tc_head2_kernel<72, 96, true>, 2 * 128 + 32);
            }
            if (half == 0) return run(tc_head2w_kernel<72, 80, false>, 2 * 256 + 32);
            return run(tc_head2w_kernel<72, 96, true>, 2 * 256 + 32);
        }
    }
    static const bool no_banded_heads = getenv("YFV2_HEADS_NOBANDS") != nullptr;
    if (W % 2 == 0 && W / 2 <= 256 && !force_band && !force_g4 && !no_banded_heads) {
        // large maps (640x640 input: 40x40): the same streamed pixel-pair kernel over bands of rows of one image
        HeadArgs a{};
        a.N = N; a.nout = 72; a.imgs = 1;
        for (int b = 0; b < 2; ++b) { a.wdw[b] = wdw[b]; a.wpw[b] = wpw[b]; }
        if (half == 0) { a.in[0] = sIn; a.in[1] = sIn; a.out[0] = tcls; a.out[1] = treg; }
        else {
            a.in[0] = tcls; a.in[1] = treg;
            a.dstA[0] = obj; a.dstB[0] = cls; a.split[0] = A; a.M[0] = A + C;
            a.dstA[1] = reg; a.dstB[1] = reg; a.split[1] = 4 * A; a.M[1] = 4 * A;
        }
        a.st2_ok = heads_st2_ok(a, half, W);
        const int Wp = W / 2;
        int BR = 256 / Wp;
        if (BR > H) BR = H;
        const int bands = (H + BR - 1) / BR;
        BR = (H + bands - 1) / bands;                           // equalise the bands
        a.band_rows = BR; a.bands = (H + BR - 1) / BR;
        const size_t PSb = (size_t)(BR + 4) * sIn.Ws;
        const size_t wfl = (size_t)(2 * np * 72 + 2 * np) + 72 * 28;
        const size_t bytes = (wfl + kHeadBufs * 8 * PSb + 4) * sizeof(float);
        if (bytes <= kSmemCap - 1024) {
            const int items = 2 * N * a.bands;
            auto run = [&](auto kern) -> int {
                TRYL(set_smem_attr(kern, bytes));
                YFV2_CUDA(launch_k(kern, min(items, sm_count()), 2 * 256 + 32, bytes, s, pdl_take(), a));
                YFV2_LAUNCH_CHECK();
                return YFV2_OK;
            };
            if (half == 0) return run(tc_head2w_kernel<72, 80, false>);
            return run(tc_head2w_kernel<72, 96, true>);
        }
    }
    if (H * W <= 512 && !force_band) {
        constexpr int G = 4;
        HeadArgs a{};
        a.N = N; a.nout = 72;
        for (int b = 0; b < 2; ++b) { a.wdw[b] = wdw[b]; a.wpw[b] = wpw[b]; }
        if (half == 0) { a.in[0] = sIn; a.in[1] = sIn; a.out[0] = tcls; a.out[1] = treg; }
        else {
            a.in[0] = tcls; a.in[1] = treg;
            a.dstA[0] = obj; a.dstB[0] = cls; a.split[0] = A; a.M[0] = A + C;
            a.dstA[1] = reg; a.dstB[1] = reg; a.split[1] = 4 * A; a.M[1] = 4 * A;
        }
        const size_t PS = (size_t)(H + 4) * sIn.Ws;
        const size_t wfl = (size_t)(2 * np * 72 + 2 * np) + 72 * 28;
        a.imgs = 1;
        while ((a.imgs + 1) * H * W <= G * 128 && a.imgs + 1 <= N &&
               (wfl + kHeadBufs * 8 * PS * (a.imgs + 1) + 4) * sizeof(float) <= kSmemCap - 1024) ++a.imgs;
        const size_t bytes = (wfl + kHeadBufs * 8 * PS * a.imgs + 4) * sizeof(float);
        if (bytes <= kSmemCap - 1024) {
            const int ngroups = (N + a.imgs - 1) / a.imgs;
            auto run = [&](auto kern) -> int {
                TRYL(set_smem_attr(kern, bytes));
                YFV2_CUDA(launch_k(kern, min(2 * ngroups, sm_count()), G * 128 + 32, bytes, s, pdl_take(), a));
                YFV2_LAUNCH_CHECK();
                return YFV2_OK;
            };
            if (half == 0) return run(tc_head_kernel<72, 80, G, false>);
            return run(tc_head_kernel<72, 96, G, true>);
        }
    }
    // ---- fallback: row bands of the whole 72-channel stack --------------------------------------------------------
    DwPwArgs a{};
    ChanTab ident;
    for (int i = 0; i < kMaxCh; ++i) ident.c[i] = (unsigned short)i;
    a.N = N; a.nbranch = 2; a.nout = 72;
    for (int b = 0; b < 2; ++b) { a.tin[b] = ident; a.tout[b] = ident; a.wdw[b] = wdw[b]; a.wpw[b] = wpw[b]; }
    if (half == 0) { a.in[0] = sIn; a.in[1] = sIn; a.out[0] = tcls; a.out[1] = treg; }
    else {
        a.in[0] = tcls; a.in[1] = treg; a.out[0] = tcls; a.out[1] = treg;
        a.dstA[0] = obj; a.dstB[0] = cls; a.split[0] = A; a.M[0] = A + C;
        a.dstA[1] = reg; a.dstB[1] = reg; a.split[1] = 4 * A; a.M[1] = 4 * A;
    }
    constexpr int G = 2;
    auto run = [&](auto kern) -> int {
        const size_t wfl = (size_t)(2 * np * 72 + 2 * np) + 72 * 28;
        dwpw_geometry(a, H, wfl + 4, (size_t)72 * sIn.Ws, 5, 1, W, G);
        const size_t bytes = (wfl + (size_t)72 * (a.TR + 4) * sIn.Ws * a.imgs + 4) * sizeof(float);
        TRYL(set_smem_attr(kern, bytes));
        const int items = ((N + a.imgs - 1) / a.imgs) * a.bandsPerImg * 2;
        YFV2_CUDA(launch_k(kern, min(items, sm_count()), G * 128, bytes, s, pdl_take(), a));
        YFV2_LAUNCH_CHECK();
        return YFV2_OK;
    };
    if (half == 0) return run(tc_dwpw_kernel<72, 80, G, 5, 1, true, false, false>);
    return run(tc_dwpw_kernel<72, 96, G, 5, 1, true, false, true>);
}

}  // namespace yfv2

/* test hook (host only): the lane map tc_launch_heads builds for an H x W map in frames of row stride WS and plane stride PS (floats)
 * with `imgs` images per item: out[256] entries (image << 16 | row << 8 | pair column), 0xFFFFFFFF = idle lane. */
extern "C" int yfv2_debug_head_lanemap(int H, int W, int WS, int imgs, long long PS, unsigned int* out) {
    if (!out || H <= 0 || W <= 0 || H > 255 || (W + 1) / 2 > 255 || imgs <= 0 || imgs * H * ((W + 1) / 2) > 256 || (WS & 1) || (PS & 1)) {
        yfv2::set_error("debug_head_lanemap: bad geometry");
        return YFV2_EINVAL;
    }
    yfv2::HeadArgs a{};
    yfv2::build_lanemap(a, H, W, WS, imgs, (size_t)PS);
    memcpy(out, a.lanemap, sizeof(a.lanemap));
    return YFV2_OK;
}
