// K0 — stem: conv3x3 s2 p1 (3->24, no bias) + BN + ReLU + maxpool3x3 s2 p1, one pass.
// Reference: model/backbone/shufflenetv2.py:74-80,103-104.  Input NCHW fp32 (or uint8 with the
// `/255.0` of utils/utils.py:368 fused into the load), output 24 planes at H/4 x W/4.
//
// One CTA = one 8x16 output tile of one image.  It stages the 35x67x3 input patch in shared memory,
// produces the 17x33 conv positions the tile's pool windows touch (8 output channels per pass, three
// passes), and max-pools straight out of shared memory.  Conv positions outside the conv output are
// -inf so they never win a pool window (PyTorch pads max_pool2d with -inf).
#include "common.cuh"

namespace yfv2 {

namespace {
constexpr int OT_H = 8, OT_W = 16;               // output tile (stride-4 grid)
constexpr int CT_H = 2 * OT_H + 1, CT_W = 2 * OT_W + 1;   // 17 x 33 conv positions
constexpr int IT_H = 2 * CT_H + 1, IT_W = 2 * CT_W + 1;   // 35 x 67 input pixels
constexpr int IT_WS = 67;                         // row stride (odd: conflict-free stride-2 reads)
constexpr int NPOS = CT_H * CT_W;                 // 561
constexpr int STEM_THREADS = 288;                 // 2 rounds cover 561 positions
constexpr int CG = 8;                             // channels per pass

template <bool U8>
__global__ void __launch_bounds__(STEM_THREADS)
stem_kernel(const void* __restrict__ xin, Planes out, const float* __restrict__ wpack, int H, int W, int tilesX) {
    __shared__ float s_in[3 * IT_H * IT_WS];
    __shared__ float s_conv[CG * NPOS];
    __shared__ __align__(16) float s_w[kStemPackFloats];

    const int n = blockIdx.y;
    const int ty = blockIdx.x / tilesX, tx = blockIdx.x % tilesX;
    const int oy0 = ty * OT_H, ox0 = tx * OT_W;
    const int cy0 = 2 * oy0 - 1, cx0 = 2 * ox0 - 1;
    const int iy0 = 2 * cy0 - 1, ix0 = 2 * cx0 - 1;
    const int HC = H / 2, WC = W / 2, HO = H / 4, WO = W / 4;
    const int tid = threadIdx.x;

    copy_to_smem(s_w, wpack, kStemPackFloats);
    // input patch: fp32 goes through cp.async (all loads in flight at once); uint8 is converted on the fly
    // in batches of 8 independent loads per thread
    if (!U8) {
        for (int i = tid; i < 3 * IT_H * IT_W; i += STEM_THREADS) {
            const int c = i / (IT_H * IT_W);
            const int rem = i - c * (IT_H * IT_W);
            const int r = rem / IT_W, q = rem - r * IT_W;
            const int iy = iy0 + r, ix = ix0 + q;
            float* dst = &s_in[(c * IT_H + r) * IT_WS + q];
            if (iy >= 0 && iy < H && ix >= 0 && ix < W)
                cp_async4(dst, reinterpret_cast<const float*>(xin) + ((((size_t)n * 3 + c) * H + iy) * W + ix));
            else
                *dst = 0.f;
        }
        cp_async_wait_all();
    } else {
        constexpr int TOT = 3 * IT_H * IT_W, UNR = 8;
        for (int i0 = tid; i0 < TOT; i0 += STEM_THREADS * UNR) {
            uint8_t v[UNR];
            int sidx[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int i = i0 + u * STEM_THREADS;
                v[u] = 0; sidx[u] = -1;
                if (i < TOT) {
                    const int c = i / (IT_H * IT_W);
                    const int rem = i - c * (IT_H * IT_W);
                    const int r = rem / IT_W, q = rem - r * IT_W;
                    const int iy = iy0 + r, ix = ix0 + q;
                    sidx[u] = (c * IT_H + r) * IT_WS + q;
                    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
                        v[u] = __ldg(reinterpret_cast<const uint8_t*>(xin) + ((((size_t)n * 3 + c) * H + iy) * W + ix));
                }
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u)
                if (sidx[u] >= 0) s_in[sidx[u]] = __fdiv_rn((float)v[u], 255.0f);
        }
    }
    __syncthreads();

    const float* s_scale = s_w + 27 * 24;
    const float* s_shift = s_scale + 24;

    for (int g = 0; g < 24 / CG; ++g) {
        for (int pos = tid; pos < NPOS; pos += STEM_THREADS) {
            const int ly = pos / CT_W, lx = pos - ly * CT_W;
            const int cy = cy0 + ly, cx = cx0 + lx;
            float res[CG];
            if (cy >= 0 && cy < HC && cx >= 0 && cx < WC) {
                float acc[CG];
#pragma unroll
                for (int j = 0; j < CG; ++j) acc[j] = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const float v = s_in[(c * IT_H + 2 * ly + ky) * IT_WS + 2 * lx + kx];
                            const float* wr = s_w + (c * 9 + ky * 3 + kx) * 24 + g * CG;
                            const float4 w0 = *reinterpret_cast<const float4*>(wr);
                            const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
                            acc[0] = fmaf(w0.x, v, acc[0]); acc[1] = fmaf(w0.y, v, acc[1]);
                            acc[2] = fmaf(w0.z, v, acc[2]); acc[3] = fmaf(w0.w, v, acc[3]);
                            acc[4] = fmaf(w1.x, v, acc[4]); acc[5] = fmaf(w1.y, v, acc[5]);
                            acc[6] = fmaf(w1.z, v, acc[6]); acc[7] = fmaf(w1.w, v, acc[7]);
                        }
#pragma unroll
                for (int j = 0; j < CG; ++j)
                    res[j] = fmaxf(fmaf(acc[j], s_scale[g * CG + j], s_shift[g * CG + j]), 0.f);
            } else {
#pragma unroll
                for (int j = 0; j < CG; ++j) res[j] = -INFINITY;
            }
#pragma unroll
            for (int j = 0; j < CG; ++j) s_conv[j * NPOS + pos] = res[j];
        }
        __syncthreads();
        for (int i = tid; i < CG * OT_H * OT_W; i += STEM_THREADS) {
            const int j = i / (OT_H * OT_W);
            const int rem = i - j * (OT_H * OT_W);
            const int oyl = rem / OT_W, oxl = rem - oyl * OT_W;
            const int oy = oy0 + oyl, ox = ox0 + oxl;
            if (oy < HO && ox < WO) {
                const float* cv = s_conv + j * NPOS + (2 * oyl) * CT_W + 2 * oxl;
                float m = cv[0];
                m = fmaxf(m, cv[1]); m = fmaxf(m, cv[2]);
                m = fmaxf(m, cv[CT_W]); m = fmaxf(m, cv[CT_W + 1]); m = fmaxf(m, cv[CT_W + 2]);
                m = fmaxf(m, cv[2 * CT_W]); m = fmaxf(m, cv[2 * CT_W + 1]); m = fmaxf(m, cv[2 * CT_W + 2]);
                plane_ptr(out, n, g * CG + j)[out.org + oy * out.Ws + ox] = m;
            }
        }
        __syncthreads();
    }
}
}  // namespace

int launch_stem(const StemArgs& a, cudaStream_t s) {
    const int HO = a.H / 4, WO = a.W / 4;
    const int tilesX = (WO + OT_W - 1) / OT_W, tilesY = (HO + OT_H - 1) / OT_H;
    dim3 grid(tilesX * tilesY, a.N);
    if (a.is_u8) stem_kernel<true><<<grid, STEM_THREADS, 0, s>>>(a.x, a.out, a.wpack, a.H, a.W, tilesX);
    else         stem_kernel<false><<<grid, STEM_THREADS, 0, s>>>(a.x, a.out, a.wpack, a.H, a.W, tilesX);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}

}  // namespace yfv2
