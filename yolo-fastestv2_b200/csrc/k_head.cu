// K4 + K5 — the four DWConvblock heads and the three shared output convs.
// Reference: model/fpn.py:5-29 (DWConvblock: dw5x5 p2 +BN+ReLU -> pw+BN -> dw5x5+BN+ReLU -> pw+BN),
//            model/detector.py:17-19,25-31 (output_{reg,obj,cls}_layers, 1x1 with bias, shared by levels;
//            obj and cls both read the cls head, fpn.py:54,61).
//
// Two launches per pyramid level, each covering both heads (blockIdx.y: 0 = cls head, 1 = reg head):
//   half A:  T = BN(pw(ReLU(BN(dw5x5(S)))))                              -> 72 scratch planes
//   half B:  F = BN(pw(ReLU(BN(dw5x5(T)))));  preds = out_conv(F) + bias -> dense NCHW outputs
// The depthwise result is consumed from registers by the pointwise accumulation; in half B the 72
// head features go through shared memory once so the output convs can be spread over all threads.
#include "common.cuh"

namespace yfv2 {
namespace {
constexpr int NT = 512;
constexpr int CH = 72;
constexpr int HALF_FLOATS = dw5_pack_floats(CH) + pw_pack_floats(CH, CH);   // DW5 | PW

struct HeadIO {
    Planes in[2];        // per head
    Planes out[2];       // half A only
    const float* w[2];   // per head: this half's DW5|PW pack
    const float* wout[2];// half B: out-conv pack per head (PW layout, scale ignored, shift = bias)
    float* dstA[2];      // half B: first destination  (obj | reg)
    float* dstB[2];      // half B: second destination (cls | unused)
    int split[2];        // outputs [0,split) -> dstA, [split,M) -> dstB
    int M[2];
};

template <int NSPLIT, bool FINAL>
__global__ void __launch_bounds__(NT)
head_kernel(HeadIO io, ChanTab ident, int TR, int tilesPerImg) {
    extern __shared__ __align__(16) float smem[];
    constexpr int NS = CH / NSPLIT;
    const int hd = blockIdx.y;
    const Planes Pin = io.in[hd];
    const int W = Pin.W, H = Pin.H, WS = W + 4;
    const int RS = (TR + 4) * WS;
    float* X = smem;
    float* wdw = X + ((CH * RS + 3) & ~3);
    float* wpw = wdw + dw5_pack_floats(CH);
    float* wo = wpw + pw_pack_floats(CH, CH);
    const int M = io.M[hd], Mp = round4(M);
    copy_to_smem(wdw, io.w[hd], HALF_FLOATS);
    if (FINAL) copy_to_smem(wo, io.wout[hd], pw_pack_floats(CH, M));

    const int tile = blockIdx.x;
    const int n = tile / tilesPerImg;
    const int r0 = (tile - n * tilesPerImg) * TR;
    const int rows = min(TR, H - r0);
    stage_rows<CH, 2, NT>(X, RS, WS, Pin, ident, n, r0 - 2, rows + 4);
    __syncthreads();

    const int npix = rows * W;
    const int items = npix * NSPLIT;
    const float* scale = wpw + CH * CH;
    const float* shift = scale + CH;
    // every thread handles at most one item per round; FINAL needs the barrier between rounds
    for (int base = 0; base < items; base += NT) {
        const int q = base + threadIdx.x;
        const bool active = q < items;
        float acc[1][NS];
        int h = 0, pix = 0, orow = 0, ox = 0;
        if (active) {
            h = q / npix;
            pix = q - h * npix;
            orow = pix / W;
            ox = pix - orow * W;
            const float* win = X + orow * WS + ox;       // top-left of the 5x5 window (pad 2 both ways)
#pragma unroll
            for (int j = 0; j < NS; ++j) acc[0][j] = 0.f;
#pragma unroll 2
            for (int k = 0; k < CH; ++k) {
                const float4* wk4 = reinterpret_cast<const float4*>(wdw + k * 28);
                float wk[28];
#pragma unroll
                for (int t = 0; t < 7; ++t) {
                    const float4 w4 = wk4[t];
                    wk[4 * t] = w4.x; wk[4 * t + 1] = w4.y; wk[4 * t + 2] = w4.z; wk[4 * t + 3] = w4.w;
                }
                const float* xk = win + k * RS;
                float d = 0.f;
#pragma unroll
                for (int dy = 0; dy < 5; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 5; ++dx) d = fmaf(wk[dy * 5 + dx], xk[dy * WS + dx], d);
                const float dv[1] = {fmaxf(fmaf(d, wk[25], wk[26]), 0.f)};      // BN + ReLU (fpn.py:13-14,20-21)
                fma_row<NS, 1>(wpw + k * CH + h * NS, dv, acc);
            }
        }
        if (!FINAL) {
            if (active) {
                const long long o = (long long)(r0 + orow) * W + ox;
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const int nn = h * NS + j;
                    plane_ptr(io.out[hd], n, nn)[o] = fmaf(acc[0][j], scale[nn], shift[nn]);   // BN, no ReLU
                }
            }
        } else {
            // features -> shared (planes [CH][npix] laid over X once every thread is done reading it)
            __syncthreads();
            if (active) {
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const int nn = h * NS + j;
                    X[nn * npix + pix] = fmaf(acc[0][j], scale[nn], shift[nn]);
                }
            }
            __syncthreads();
        }
    }
    if (FINAL) {
        // NOTE: the loop above runs exactly one round whenever FINAL (launcher guarantees items <= NT),
        // so X now holds the complete feature tile.
        const float* bias = wo + CH * Mp + Mp;
        const int HW = H * W;
        const int chunks = Mp / 4;
        for (int it = threadIdx.x; it < npix * chunks; it += NT) {
            const int m4 = it / npix, pix = it - m4 * npix;
            float4 a = *reinterpret_cast<const float4*>(bias + 4 * m4);
#pragma unroll 8
            for (int k = 0; k < CH; ++k) {
                const float f = X[k * npix + pix];
                const float4 w = *reinterpret_cast<const float4*>(wo + k * Mp + 4 * m4);
                a.x = fmaf(w.x, f, a.x); a.y = fmaf(w.y, f, a.y); a.z = fmaf(w.z, f, a.z); a.w = fmaf(w.w, f, a.w);
            }
            const float r[4] = {a.x, a.y, a.z, a.w};
            const long long p = (long long)r0 * W + pix;
            const int split = io.split[hd];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = 4 * m4 + e;
                if (m < split) io.dstA[hd][((long long)n * split + m) * HW + p] = r[e];
                else if (m < M) io.dstB[hd][((long long)n * (M - split) + (m - split)) * HW + p] = r[e];
            }
        }
    }
}

template <int NSPLIT, bool FINAL>
int run_half(const HeadIO& io, const ChanTab& ident, int N, int TR, size_t bytes, cudaStream_t s) {
    auto kern = head_kernel<NSPLIT, FINAL>;
    YFV2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const int H = io.in[0].H;
    const int tilesPerImg = (H + TR - 1) / TR;
    dim3 grid(tilesPerImg * N, 2);
    kern<<<grid, NT, bytes, s>>>(io, ident, TR, tilesPerImg);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}
}  // namespace

size_t head_pack_floats() { return 2 * (size_t)HALF_FLOATS; }

int launch_heads(const HeadArgs& a, int half, cudaStream_t s) {
    const int H = a.s.H, W = a.s.W;
    const int Moc = a.A + a.C, Mreg = 4 * a.A;
    const size_t wout_max = (size_t)pw_pack_floats(CH, Moc > Mreg ? Moc : Mreg);
    // band height: NSPLIT * rows * W work items must fit one round of NT threads
    int nsplit = 2;
    int TR = H;
    while (TR > 1 && TR * W * nsplit > NT) --TR;
    if (TR * W * 3 <= NT) nsplit = 3;
    auto bytes = [&](int tr, bool fin) {
        return (size_t)(((CH * (tr + 4) * (W + 4) + 3) & ~3) + HALF_FLOATS + (fin ? wout_max : 0)) * sizeof(float);
    };
    while (TR > 1 && bytes(TR, true) > kSmemCap) --TR;
    if (W * nsplit > NT || bytes(TR, true) > kSmemCap) {
        set_error("launch_heads: feature map %dx%d too wide for the head kernel", H, W);
        return YFV2_EUNSUPPORTED;
    }
    ChanTab ident;
    for (int i = 0; i < kMaxCh; ++i) ident.c[i] = (unsigned short)i;

    HeadIO io{};
    io.in[0] = a.s; io.in[1] = a.s;
    io.out[0] = a.t_cls; io.out[1] = a.t_reg;
    io.w[0] = a.w_cls; io.w[1] = a.w_reg;
    if (half == 0)
        return nsplit == 3 ? run_half<3, false>(io, ident, a.N, TR, bytes(TR, false), s)
                           : run_half<2, false>(io, ident, a.N, TR, bytes(TR, false), s);

    io.in[0] = a.t_cls; io.in[1] = a.t_reg;
    io.w[0] = a.w_cls + HALF_FLOATS; io.w[1] = a.w_reg + HALF_FLOATS;
    io.wout[0] = a.w_out_oc; io.wout[1] = a.w_out_reg;
    io.dstA[0] = a.obj; io.dstB[0] = a.cls; io.split[0] = a.A; io.M[0] = Moc;
    io.dstA[1] = a.reg; io.dstB[1] = a.reg; io.split[1] = Mreg; io.M[1] = Mreg;
    return nsplit == 3 ? run_half<3, true>(io, ident, a.N, TR, bytes(TR, true), s)
                       : run_half<2, true>(io, ident, a.N, TR, bytes(TR, true), s);
}

}  // namespace yfv2
