// K1 / K2 — fused ShuffleNetV2 blocks (reference model/backbone/shufflenetv2.py:5-63), eval mode.
//
// stride 1 (K2):  out_main = ReLU(BN(pw2( BN(dw3x3( ReLU(BN(pw1(x_odd))) )) )))   — one kernel.
//                 The even ("passthrough") channels are not touched: the plan just keeps their plane ids.
// stride 2 (K1):  proj = ReLU(BN(pw( BN(dw3x3s2(x)) )));  main as above with a stride-2 depthwise.
//
// One CTA owns a band of output rows of one image.  It stages the band (+1 halo row each side, zero
// padded columns) of the K input planes in shared memory, runs pw1 IN PLACE on every staged pixel (a
// thread owns a pixel column, so in-place is race free; with the output channels split over NSPLIT
// threads the stores are separated from the loads by a barrier), zeroes nothing twice (out-of-image
// rows stay 0, which is exactly the depthwise zero padding of the *pw1 output*), then walks the K
// channels once more: depthwise 3x3 in registers -> BN -> immediately accumulated into the pw2 outputs.
// The depthwise result never touches memory.
#include "common.cuh"

namespace yfv2 {

namespace {
constexpr int NT = 256;

// pack offsets (floats)
__host__ __device__ constexpr int pwf(int K) { return pw_pack_floats(K, K); }

// pw1 in place over the staged rows that are inside the image.
//   items = nrows * (W/PPT) pixel groups x NSPLIT output slices.
template <int K, int PPT, int NSPLIT>
__device__ __forceinline__ void pw_inplace(float* __restrict__ X, int RS, int WS, int W, int H, int gr0, int nrows,
                                           const float* __restrict__ wpw) {
    constexpr int NS = K / NSPLIT;
    const int groups = W / PPT;
    const int cnt = nrows * groups;
    const float* scale = wpw + K * K;
    const float* shift = scale + K;
    // All NSPLIT output slices of a pixel group run in the SAME round: the stores of a round overwrite the
    // inputs of exactly the pixel groups that round has finished reading.
    constexpr int PPR = NT / NSPLIT;
    const int h = threadIdx.x / PPR;
    for (int base = 0; base < cnt; base += PPR) {
        const int r = base + (threadIdx.x - h * PPR);
        bool active = r < cnt && h < NSPLIT;
        int rr = 0, x0 = 0;
        if (active) {
            rr = r / groups;
            x0 = (r - rr * groups) * PPT;
            const int gr = gr0 + rr;
            active = (gr >= 0 && gr < H);
        }
        float acc[PPT][NS];
        float* px = X + rr * WS + 1 + x0;
        if (active) {
#pragma unroll
            for (int i = 0; i < PPT; ++i)
#pragma unroll
                for (int j = 0; j < NS; ++j) acc[i][j] = 0.f;
#pragma unroll 2
            for (int k = 0; k < K; ++k) {
                float xv[PPT];
#pragma unroll
                for (int i = 0; i < PPT; ++i) xv[i] = px[k * RS + i];
                fma_row<NS, PPT>(wpw + k * K + h * NS, xv, acc);
            }
        }
        if (NSPLIT > 1) __syncthreads();
        if (active) {
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const float s = scale[h * NS + j], b = shift[h * NS + j];
#pragma unroll
                for (int i = 0; i < PPT; ++i) px[(h * NS + j) * RS + i] = fmaxf(fmaf(acc[i][j], s, b), 0.f);
            }
        }
        if (NSPLIT > 1) __syncthreads();
    }
}

// depthwise 3x3 (stride S) + BN from staged planes, fused into a K->K pointwise + BN + ReLU; stores to
// the output planes tout[ch_off + n].  Local staged row of output row `orow` centre: S*orow + 1 when the
// staged band starts one input row above the band (true for both strides here).
template <int K, int PPT, int NSPLIT, int S>
__device__ __forceinline__ void dw_pw_store(const float* __restrict__ X, int RS, int WS, int Wout, int rows,
                                            const float* __restrict__ wdw, const float* __restrict__ wpw,
                                            const Planes& Pout, const ChanTab& tout, int ch_off, int n, int r0) {
    constexpr int NS = K / NSPLIT;
    constexpr int WIN = S * (PPT - 1) + 3;     // window columns covering PPT adjacent outputs
    const int groups = Wout / PPT;
    const int cnt = rows * groups;
    const int items = cnt * NSPLIT;
    const float* scale = wpw + K * K;
    const float* shift = scale + K;
    for (int q = threadIdx.x; q < items; q += NT) {
        const int h = q / cnt;
        const int r = q - h * cnt;
        const int orow = r / groups;
        const int ox0 = (r - orow * groups) * PPT;
        // top-left of the window in staged coordinates: row S*orow, col S*ox0 (pad column included)
        const float* win = X + (S * orow) * WS + S * ox0;
        float acc[PPT][NS];
#pragma unroll
        for (int i = 0; i < PPT; ++i)
#pragma unroll
            for (int j = 0; j < NS; ++j) acc[i][j] = 0.f;
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            const float4 wa = *reinterpret_cast<const float4*>(wdw + k * 12);
            const float4 wb = *reinterpret_cast<const float4*>(wdw + k * 12 + 4);
            const float4 wc = *reinterpret_cast<const float4*>(wdw + k * 12 + 8);
            const float w9[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x};
            float d[PPT];
#pragma unroll
            for (int i = 0; i < PPT; ++i) d[i] = 0.f;
            const float* wk = win + k * RS;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                float v[WIN];
#pragma unroll
                for (int c = 0; c < WIN; ++c) v[c] = wk[dy * WS + c];
#pragma unroll
                for (int i = 0; i < PPT; ++i)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) d[i] = fmaf(w9[dy * 3 + dx], v[S * i + dx], d[i]);
            }
#pragma unroll
            for (int i = 0; i < PPT; ++i) d[i] = fmaf(d[i], wc.y, wc.z);     // BN, no ReLU (shufflenetv2.py:25-26)
            fma_row<NS, PPT>(wpw + k * K + h * NS, d, acc);
        }
        const long long o = (long long)(r0 + orow) * Wout + ox0;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const int nn = h * NS + j;
            float* dst = plane_ptr(Pout, n, tout.c[ch_off + nn]) + o;
            const float s = scale[nn], b = shift[nn];
#pragma unroll
            for (int i = 0; i < PPT; ++i) dst[i] = fmaxf(fmaf(acc[i][j], s, b), 0.f);
        }
    }
}

// ---- stride 1 -----------------------------------------------------------------------------------
// wpack: PW1 | DW | PW2
template <int K, int PPT, int NSPLIT>
__global__ void __launch_bounds__(NT)
shuffle_s1_kernel(Planes P, ChanTab tin, ChanTab tout, const float* __restrict__ wpack, int TR, int tilesPerImg, int total) {
    extern __shared__ __align__(16) float smem[];
    const int W = P.W, H = P.H, WS = W + 2;
    const int RS = (TR + 2) * WS;
    float* X = smem;
    float* w1 = X + ((K * RS + 3) & ~3);
    float* wd = w1 + pwf(K);
    float* w2 = wd + dw3_pack_floats(K);
    copy_to_smem(w1, wpack, 2 * pwf(K) + dw3_pack_floats(K));

    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int n = tile / tilesPerImg;
        const int r0 = (tile - n * tilesPerImg) * TR;
        const int rows = min(TR, H - r0);
        __syncthreads();                                   // previous tile fully consumed; weights landed
        stage_rows<K, 1, NT>(X, RS, WS, P, tin, n, r0 - 1, rows + 2);
        __syncthreads();
        pw_inplace<K, PPT, NSPLIT>(X, RS, WS, W, H, r0 - 1, rows + 2, w1);
        __syncthreads();
        dw_pw_store<K, PPT, NSPLIT, 1>(X, RS, WS, W, rows, wd, w2, P, tout, 0, n, r0);
    }
}

// ---- stride 2 -----------------------------------------------------------------------------------
// wpack: DWp | PWp | PW1 | DW | PW2.  PW weights stream through ONE shared buffer when !RESIDENT.
template <int K, int PPT1, int PPT2, int NSPLIT, bool RESIDENT>
__global__ void __launch_bounds__(NT)
shuffle_s2_kernel(Planes Pin, Planes Pout, ChanTab tin, ChanTab tout, const float* __restrict__ wpack, int TR,
                  int tilesPerImg, int total) {
    extern __shared__ __align__(16) float smem[];
    const int Win = Pin.W, Hin = Pin.H, WS = Win + 2;
    const int Wout = Pout.W, Hout = Pout.H;
    const int RS = (2 * TR + 1) * WS;
    float* X = smem;
    float* wdp = X + ((K * RS + 3) & ~3);
    float* wdm = wdp + dw3_pack_floats(K);
    float* wpa = wdm + dw3_pack_floats(K);                       // PWp (or the streaming buffer)
    float* wpb = RESIDENT ? wpa + pwf(K) : wpa;                  // PW1
    float* wpc = RESIDENT ? wpb + pwf(K) : wpa;                  // PW2
    const float* g_dwp = wpack;
    const float* g_pwp = g_dwp + dw3_pack_floats(K);
    const float* g_pw1 = g_pwp + pwf(K);
    const float* g_dwm = g_pw1 + pwf(K);
    const float* g_pw2 = g_dwm + dw3_pack_floats(K);
    copy_to_smem(wdp, g_dwp, dw3_pack_floats(K));
    copy_to_smem(wdm, g_dwm, dw3_pack_floats(K));
    if (RESIDENT) {
        copy_to_smem(wpa, g_pwp, pwf(K));
        copy_to_smem(wpb, g_pw1, pwf(K));
        copy_to_smem(wpc, g_pw2, pwf(K));
    }

    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int n = tile / tilesPerImg;
        const int r0 = (tile - n * tilesPerImg) * TR;
        const int rows = min(TR, Hout - r0);
        const int gr0 = 2 * r0 - 1, nrows = 2 * rows + 1;
        __syncthreads();
        stage_rows<K, 1, NT>(X, RS, WS, Pin, tin, n, gr0, nrows);
        if (!RESIDENT) copy_to_smem(wpa, g_pwp, pwf(K));
        __syncthreads();
        // branch_proj on the raw input (shufflenetv2.py:34-44)
        dw_pw_store<K, PPT2, NSPLIT, 2>(X, RS, WS, Wout, rows, wdp, wpa, Pout, tout, 0, n, r0);
        __syncthreads();
        if (!RESIDENT) { copy_to_smem(wpb, g_pw1, pwf(K)); __syncthreads(); }
        pw_inplace<K, PPT1, NSPLIT>(X, RS, WS, Win, Hin, gr0, nrows, wpb);
        __syncthreads();
        if (!RESIDENT) { copy_to_smem(wpc, g_pw2, pwf(K)); __syncthreads(); }
        dw_pw_store<K, PPT2, NSPLIT, 2>(X, RS, WS, Wout, rows, wdm, wpc, Pout, tout, K, n, r0);
    }
}

template <typename Kern>
int set_smem(Kern kern, size_t bytes) {
    YFV2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return YFV2_OK;
}

template <int K, int PPT, int NSPLIT>
int run_s1(const ShuffleArgs& a, cudaStream_t s) {
    const int W = a.in.W, H = a.in.H;
    const size_t wfl = 2 * pwf(K) + dw3_pack_floats(K);
    // rows per band: largest band that keeps two CTAs per SM when possible
    int TR = H;
    auto bytes = [&](int tr) { return (size_t)(((K * (tr + 2) * (W + 2) + 3) & ~3) + wfl) * sizeof(float); };
    while (TR > 1 && bytes(TR) > 110 * 1024) TR = (TR + 1) / 2;
    if (bytes(TR) > kSmemCap) { set_error("shuffle_s1: tile does not fit shared memory (K=%d W=%d)", K, W); return YFV2_EUNSUPPORTED; }
    const int tilesPerImg = (H + TR - 1) / TR;
    const int total = tilesPerImg * a.N;
    auto kern = shuffle_s1_kernel<K, PPT, NSPLIT>;
    int rc = set_smem(kern, bytes(TR));
    if (rc) return rc;
    const int grid = min(total, sm_count() * 4);
    kern<<<grid, NT, bytes(TR), s>>>(a.in, a.tin, a.tout, a.wpack, TR, tilesPerImg, total);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}

template <int K, int PPT1, int PPT2, int NSPLIT, bool RESIDENT>
int run_s2(const ShuffleArgs& a, cudaStream_t s) {
    const int Win = a.in.W, Hout = a.out.H;
    const size_t wfl = 2 * dw3_pack_floats(K) + (RESIDENT ? 3 : 1) * pwf(K);
    auto bytes = [&](int tr) { return (size_t)(((K * (2 * tr + 1) * (Win + 2) + 3) & ~3) + wfl) * sizeof(float); };
    int TR = Hout;
    while (TR > 1 && bytes(TR) > 112 * 1024) TR = TR - 1;
    if (bytes(TR) > kSmemCap) { set_error("shuffle_s2: tile does not fit shared memory (K=%d W=%d)", K, Win); return YFV2_EUNSUPPORTED; }
    const int tilesPerImg = (Hout + TR - 1) / TR;
    const int total = tilesPerImg * a.N;
    auto kern = shuffle_s2_kernel<K, PPT1, PPT2, NSPLIT, RESIDENT>;
    int rc = set_smem(kern, bytes(TR));
    if (rc) return rc;
    const int grid = min(total, sm_count() * 4);
    kern<<<grid, NT, bytes(TR), s>>>(a.in, a.out, a.tin, a.tout, a.wpack, TR, tilesPerImg, total);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}
}  // namespace

size_t shuffle_pack_floats(int K, int stride) {
    return stride == 1 ? (size_t)2 * pwf(K) + dw3_pack_floats(K) : (size_t)3 * pwf(K) + 2 * dw3_pack_floats(K);
}

int launch_shuffle(const ShuffleArgs& a, cudaStream_t s) {
    if (a.stride == 1) {
        const bool even = (a.in.W % 2) == 0;
        switch (a.K) {
            case 24: return even ? run_s1<24, 2, 1>(a, s) : run_s1<24, 1, 1>(a, s);
            case 48: return even ? run_s1<48, 2, 1>(a, s) : run_s1<48, 1, 1>(a, s);
            case 96: return run_s1<96, 1, 2>(a, s);
        }
    } else {
        const bool even = (a.out.W % 2) == 0;   // input width is always even
        switch (a.K) {
            case 24: return even ? run_s2<24, 2, 2, 1, true>(a, s) : run_s2<24, 2, 1, 1, true>(a, s);
            case 48: return even ? run_s2<48, 2, 2, 1, true>(a, s) : run_s2<48, 2, 1, 1, true>(a, s);
            case 96: return run_s2<96, 1, 1, 2, false>(a, s);
        }
    }
    set_error("launch_shuffle: unsupported K=%d stride=%d", a.K, a.stride);
    return YFV2_EUNSUPPORTED;
}

}  // namespace yfv2
