// Training-side operators: train-mode forward (batch-statistics BatchNorm) and the backward of every
// op of the network, as plain dense-NCHW fp32 kernels behind the C ABI.  The Python mirror composes them with
// torch.autograd.Function objects (yolo-fastestv2_b200/model/train_ops.py), so autograd's graph does the bookkeeping and
// the arithmetic is ours.  Round-1 goal here is correctness against the reference's autograd (tests/test_train_gpu.py);
// these are straightforward FFMA kernels, not yet the fused tcgen05 path of the inference engine.
//
//   conv1x1   fwd / dgrad / wgrad (+bias)      one generic strided batched GEMM (64x64x16 tiles, 4x4 micro-tiles)
//   dwconv    fwd / dgrad / wgrad              3x3 or 5x5, stride 1 or 2, pad k/2      (shufflenetv2.py:25,36; fpn.py:12,19)
//   stem conv fwd / wgrad                      dense 3x3 s2 p1, 3->24                   (shufflenetv2.py:75)
//   batchnorm train fwd / bwd (+ReLU)          batch statistics, running-stat update    (nn.BatchNorm2d, momentum .1, eps 1e-5)
//   maxpool 3x3 s2 p1 fwd / bwd, nearest 2x upsample fwd / bwd                           (shufflenetv2.py:80; fpn.py:57)
#include "common.cuh"

namespace yfv2 {
namespace {

// ---------------------------------------------------------------------------------------------------------------------
// Generic batched GEMM with arbitrary strides:  C[b](i,j) (+)= sum_k A[b](i,k) * B[b](k,j)
// ---------------------------------------------------------------------------------------------------------------------
struct GemmArgs {
    const float* A; const float* B; float* C;
    int M, N, K, batch;
    long long sAi, sAk, sAb, sBk, sBj, sBb, sCi, sCj, sCb;
    const float* bias;      // per-row (i) bias added once (only when !atomic), may be null
    int atomic;             // atomicAdd into C (reduction over the batch dimension with sCb == 0)
};

__global__ void __launch_bounds__(256)
gemm_kernel(GemmArgs g) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int b = blockIdx.z;
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    const float* A = g.A + b * g.sAb;
    const float* B = g.B + b * g.sBb;
    float* C = g.C + b * g.sCb;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    for (int k0 = 0; k0 < g.K; k0 += 16) {
        for (int t = threadIdx.x; t < 16 * 64; t += 256) {
            const int kk = t >> 6, ii = t & 63;
            const int k = k0 + kk;
            As[kk][ii] = (k < g.K && i0 + ii < g.M) ? A[(long long)(i0 + ii) * g.sAi + (long long)k * g.sAk] : 0.f;
            Bs[kk][ii] = (k < g.K && j0 + ii < g.N) ? B[(long long)k * g.sBk + (long long)(j0 + ii) * g.sBj] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], bb[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { a[q] = As[kk][ty * 4 + q]; bb[q] = Bs[kk][tx * 4 + q]; }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[q][r] = fmaf(a[q], bb[r], acc[q][r]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = i0 + ty * 4 + q;
        if (i >= g.M) continue;
        const float bv = (g.bias && !g.atomic) ? g.bias[i] : 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = j0 + tx * 4 + r;
            if (j >= g.N) continue;
            float* dst = C + (long long)i * g.sCi + (long long)j * g.sCj;
            if (g.atomic) atomicAdd(dst, acc[q][r]);
            else *dst = acc[q][r] + bv;
        }
    }
}

int run_gemm(const GemmArgs& g, cudaStream_t s) {
    dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, g.batch);
    gemm_kernel<<<grid, 256, 0, s>>>(g);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}

// per-row sums of dy [N,M,HW] -> dbias[M]
__global__ void bias_grad_kernel(const float* __restrict__ dy, float* __restrict__ db, int N, int M, int HW) {
    const int m = blockIdx.x;
    double acc = 0.0;
    for (long long t = threadIdx.x; t < (long long)N * HW; t += blockDim.x) {
        const int n = (int)(t / HW), p = (int)(t - (long long)n * HW);
        acc += dy[((long long)n * M + m) * HW + p];
    }
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) { if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st]; __syncthreads(); }
    if (threadIdx.x == 0) db[m] = (float)red[0];
}

// ---------------------------------------------------------------------------------------------------------------------
// depthwise conv, kernel KS (3/5), stride S, pad KS/2, dense NCHW
// ---------------------------------------------------------------------------------------------------------------------
template <int KS>
__global__ void dw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int N, int C, int H, int W,
                              int Ho, int Wo, int S) {
    const long long total = (long long)N * C * Ho * Wo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho), c = (int)((i / ((long long)Wo * Ho)) % C);
        const long long n = i / ((long long)Wo * Ho * C);
        const float* xp = x + (n * C + c) * (long long)H * W;
        const float* wp = w + c * KS * KS;
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
            const int iy = oy * S - KS / 2 + ky;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
                const int ix = ox * S - KS / 2 + kx;
                if (ix >= 0 && ix < W) acc = fmaf(wp[ky * KS + kx], xp[iy * W + ix], acc);
            }
        }
        y[i] = acc;
    }
}

template <int KS>
__global__ void dw_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int N, int C, int H, int W,
                                int Ho, int Wo, int S) {
    const long long total = (long long)N * C * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % W), iy = (int)((i / W) % H), c = (int)((i / ((long long)W * H)) % C);
        const long long n = i / ((long long)W * H * C);
        const float* dp = dy + (n * C + c) * (long long)Ho * Wo;
        const float* wp = w + c * KS * KS;
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
            const int t = iy + KS / 2 - ky;            // oy*S = t
            if (t < 0 || t % S) continue;
            const int oy = t / S;
            if (oy >= Ho) continue;
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
                const int u = ix + KS / 2 - kx;
                if (u < 0 || u % S) continue;
                const int ox = u / S;
                if (ox < Wo) acc = fmaf(wp[ky * KS + kx], dp[oy * Wo + ox], acc);
            }
        }
        dx[i] = acc;
    }
}

// one CTA per (channel, slice of the batch): KS*KS partial sums per thread, block reduce, atomicAdd into dw
template <int KS>
__global__ void __launch_bounds__(256)
dw_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int N, int C, int H, int W, int Ho, int Wo,
                int S) {
    const int c = blockIdx.x;
    float acc[KS * KS];
#pragma unroll
    for (int t = 0; t < KS * KS; ++t) acc[t] = 0.f;
    for (int n = blockIdx.y; n < N; n += gridDim.y) {
        const float* xp = x + ((long long)n * C + c) * H * W;
        const float* dp = dy + ((long long)n * C + c) * Ho * Wo;
        for (int p = threadIdx.x; p < Ho * Wo; p += 256) {
            const int oy = p / Wo, ox = p - oy * Wo;
            const float d = dp[p];
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
                const int iy = oy * S - KS / 2 + ky;
                if (iy < 0 || iy >= H) continue;
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    const int ix = ox * S - KS / 2 + kx;
                    if (ix >= 0 && ix < W) acc[ky * KS + kx] = fmaf(d, xp[iy * W + ix], acc[ky * KS + kx]);
                }
            }
        }
    }
    __shared__ float red[8][KS * KS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < KS * KS; ++t) {
        float v = acc[t];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < KS * KS) {
        float v = 0.f;
        for (int wi = 0; wi < 8; ++wi) v += red[wi][threadIdx.x];
        atomicAdd(dw + c * KS * KS + threadIdx.x, v);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stem: dense conv 3x3 s2 p1, Cin=3 -> M
// ---------------------------------------------------------------------------------------------------------------------
__global__ void stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int N, int M, int H, int W) {
    const int Ho = H / 2, Wo = W / 2;
    const long long total = (long long)N * M * Ho * Wo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho), m = (int)((i / ((long long)Wo * Ho)) % M);
        const long long n = i / ((long long)Wo * Ho * M);
        float acc = 0.f;
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int iy = 2 * oy - 1 + ky;
                if (iy < 0 || iy >= H) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int ix = 2 * ox - 1 + kx;
                    if (ix >= 0 && ix < W) acc = fmaf(w[((m * 3 + c) * 3 + ky) * 3 + kx], x[((n * 3 + c) * H + iy) * W + ix], acc);
                }
            }
        y[i] = acc;
    }
}

// grid (M, slices): 27 partial sums per thread over (n, oy, ox), block reduce, atomicAdd
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int N, int M, int H, int W) {
    const int Ho = H / 2, Wo = W / 2;
    const int m = blockIdx.x;
    float acc[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) acc[t] = 0.f;
    const long long total = (long long)N * Ho * Wo;
    for (long long q = (long long)blockIdx.y * 256 + threadIdx.x; q < total; q += (long long)gridDim.y * 256) {
        const int ox = (int)(q % Wo), oy = (int)((q / Wo) % Ho);
        const long long n = q / ((long long)Wo * Ho);
        const float d = dy[((n * M + m) * Ho + oy) * Wo + ox];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int iy = 2 * oy - 1 + ky;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int ix = 2 * ox - 1 + kx;
                    if (iy >= 0 && iy < H && ix >= 0 && ix < W) acc[(c * 3 + ky) * 3 + kx] = fmaf(d, x[((n * 3 + c) * H + iy) * W + ix], acc[(c * 3 + ky) * 3 + kx]);
                }
            }
    }
    __shared__ float red[8][27];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
        float v = acc[t];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < 27) {
        float v = 0.f;
        for (int wi = 0; wi < 8; ++wi) v += red[wi][threadIdx.x];
        atomicAdd(dw + m * 27 + threadIdx.x, v);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// BatchNorm (training mode)
// ---------------------------------------------------------------------------------------------------------------------
// stats[c] = (sum, sumsq) in fp64; grid (C, slices)
__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, int N, int C, int HW) {
    const int c = blockIdx.x;
    double s1 = 0.0, s2 = 0.0;
    const long long total = (long long)N * HW;
    for (long long q = (long long)blockIdx.y * 256 + threadIdx.x; q < total; q += (long long)gridDim.y * 256) {
        const long long n = q / HW;
        const float v = x[(n * C + c) * HW + (q - n * HW)];
        s1 += v; s2 += (double)v * v;
    }
    __shared__ double r1[256], r2[256];
    r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) { r1[threadIdx.x] += r1[threadIdx.x + st]; r2[threadIdx.x] += r2[threadIdx.x + st]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { atomicAdd(stats + 2 * c, r1[0]); atomicAdd(stats + 2 * c + 1, r2[0]); }
}

// mean / invstd from the sums; running-stat update as nn.BatchNorm2d: momentum 0.1, unbiased variance
__global__ void bn_finalize_kernel(const double* __restrict__ stats, float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, int C, double count, float momentum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = stats[2 * c] / count;
    double var = stats[2 * c + 1] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)kBnEps));
    if (running_mean) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y, long long total, int C,
                                int HW, int relu) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i / HW) % C);
        float v = (x[i] - mean[c]) * invstd[c] * gamma[c] + beta[c];
        if (relu) v = fmaxf(v, 0.f);
        y[i] = v;
    }
}

// sums[c] = (sum dy, sum dy*xhat) with the ReLU mask applied (y > 0); grid (C, slices)
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy, const float* __restrict__ mean,
                     const float* __restrict__ invstd, double* __restrict__ sums, int N, int C, int HW, int relu) {
    const int c = blockIdx.x;
    double s1 = 0.0, s2 = 0.0;
    const float mu = mean[c], is = invstd[c];
    const long long total = (long long)N * HW;
    for (long long q = (long long)blockIdx.y * 256 + threadIdx.x; q < total; q += (long long)gridDim.y * 256) {
        const long long n = q / HW;
        const long long idx = (n * C + c) * HW + (q - n * HW);
        float d = dy[idx];
        if (relu && !(y[idx] > 0.f)) d = 0.f;
        s1 += d; s2 += (double)d * ((x[idx] - mu) * is);
    }
    __shared__ double r1[256], r2[256];
    r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) { r1[threadIdx.x] += r1[threadIdx.x + st]; r2[threadIdx.x] += r2[threadIdx.x + st]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { atomicAdd(sums + 2 * c, r1[0]); atomicAdd(sums + 2 * c + 1, r2[0]); }
}

// dx = gamma*invstd/m * (m*dy - sum(dy) - xhat*sum(dy*xhat));  dgamma = sum(dy*xhat), dbeta = sum(dy)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                                    const double* __restrict__ sums, float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                    long long total, int C, int HW, double count, int relu) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i / HW) % C);
        float d = dy[i];
        if (relu && !(y[i] > 0.f)) d = 0.f;
        const float xh = (x[i] - mean[c]) * invstd[c];
        const float sd = (float)(sums[2 * c] / count), sdx = (float)(sums[2 * c + 1] / count);
        dx[i] = gamma[c] * invstd[c] * (d - sd - xh * sdx);
    }
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C && blockIdx.x * blockDim.x < C) {
        // only the first ceil(C/blockDim) blocks reach here with c < C
        dgamma[c] = (float)sums[2 * c + 1];
        dbeta[c] = (float)sums[2 * c];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// maxpool 3x3 s2 p1 (with argmax) and nearest 2x upsample
// ---------------------------------------------------------------------------------------------------------------------
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int* __restrict__ idx, long long planes, int H, int W) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = planes * Ho * Wo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
        const long long pl = i / ((long long)Wo * Ho);
        const float* xp = x + pl * H * W;
        float best = -INFINITY; int bi = -1;
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = 2 * oy - 1 + ky;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox - 1 + kx;
                if (ix < 0 || ix >= W) continue;
                const float v = xp[iy * W + ix];
                if (v > best || bi < 0) { best = v; bi = iy * W + ix; }      // first maximum, as ATen's max_pool2d
            }
        }
        y[i] = best; idx[i] = bi;
    }
}
__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, const int* __restrict__ idx, float* __restrict__ dx, long long planes, int H, int W) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = planes * Ho * Wo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pl = i / ((long long)Wo * Ho);
        atomicAdd(dx + pl * H * W + idx[i], dy[i]);
    }
}
__global__ void upsample2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes, int H, int W) {
    const long long total = planes * 4 * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % (2 * W)), oy = (int)((i / (2 * W)) % (2 * H));
        const long long pl = i / ((long long)4 * H * W);
        y[i] = x[pl * H * W + (oy >> 1) * W + (ox >> 1)];
    }
}
__global__ void upsample2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long planes, int H, int W) {
    const long long total = planes * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x0 = (int)(i % W), y0 = (int)((i / W) % H);
        const long long pl = i / ((long long)H * W);
        const float* d = dy + pl * 4 * H * W + (2 * y0) * (2 * W) + 2 * x0;
        dx[i] = (d[0] + d[1]) + (d[2 * W] + d[2 * W + 1]);
    }
}

int grid_for(long long total) { long long b = (total + 255) / 256; return (int)(b < 1 ? 1 : (b > 148 * 32 ? 148 * 32 : b)); }

}  // namespace
}  // namespace yfv2

using namespace yfv2;

#define ARGCHK(cond, msg) do { if (!(cond)) { set_error("%s", msg); return YFV2_EINVAL; } } while (0)

// y[n][m][p] = sum_k w[m][k] x[n][k][p] (+ bias[m])
extern "C" YFV2_API int yfv2_op_conv1x1_fwd(const float* x, const float* w, const float* bias, float* y, int N, int K, int M, int HW, void* stream) {
    ARGCHK(x && w && y && N > 0 && K > 0 && M > 0 && HW > 0, "conv1x1_fwd: bad arguments");
    GemmArgs g{w, x, y, M, HW, K, N, K, 1, 0, HW, 1, (long long)K * HW, HW, 1, (long long)M * HW, bias, 0};
    return run_gemm(g, (cudaStream_t)stream);
}
// dx[n][k][p] = sum_m w[m][k] dy[n][m][p];  dw[m][k] = sum_{n,p} dy[n][m][p] x[n][k][p];  dbias[m] = sum dy
extern "C" YFV2_API int yfv2_op_conv1x1_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* dbias, int N, int K, int M,
                                            int HW, void* stream) {
    ARGCHK(x && w && dy && N > 0 && K > 0 && M > 0 && HW > 0, "conv1x1_bwd: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (dx) {
        GemmArgs g{w, dy, dx, K, HW, M, N, 1, K, 0, HW, 1, (long long)M * HW, HW, 1, (long long)K * HW, nullptr, 0};
        int rc = run_gemm(g, s);
        if (rc) return rc;
    }
    if (dw) {
        YFV2_CUDA(cudaMemsetAsync(dw, 0, (size_t)M * K * sizeof(float), s));
        GemmArgs g{dy, x, dw, M, K, HW, N, HW, 1, (long long)M * HW, 1, HW, (long long)K * HW, K, 1, 0, nullptr, 1};
        int rc = run_gemm(g, s);
        if (rc) return rc;
    }
    if (dbias) {
        bias_grad_kernel<<<M, 256, 0, s>>>(dy, dbias, N, M, HW);
        YFV2_LAUNCH_CHECK();
    }
    return YFV2_OK;
}

extern "C" YFV2_API int yfv2_op_dwconv_fwd(const float* x, const float* w, float* y, int N, int C, int H, int W, int ks, int stride, void* stream) {
    ARGCHK(x && w && y && (ks == 3 || ks == 5) && (stride == 1 || stride == 2), "dwconv_fwd: bad arguments");
    const int Ho = (H + 2 * (ks / 2) - ks) / stride + 1, Wo = (W + 2 * (ks / 2) - ks) / stride + 1;
    const long long total = (long long)N * C * Ho * Wo;
    if (ks == 3) dw_fwd_kernel<3><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, w, y, N, C, H, W, Ho, Wo, stride);
    else dw_fwd_kernel<5><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, w, y, N, C, H, W, Ho, Wo, stride);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}
extern "C" YFV2_API int yfv2_op_dwconv_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, int N, int C, int H, int W, int ks,
                                           int stride, void* stream) {
    ARGCHK(x && w && dy && (ks == 3 || ks == 5) && (stride == 1 || stride == 2), "dwconv_bwd: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const int Ho = (H + 2 * (ks / 2) - ks) / stride + 1, Wo = (W + 2 * (ks / 2) - ks) / stride + 1;
    if (dx) {
        const long long total = (long long)N * C * H * W;
        if (ks == 3) dw_dgrad_kernel<3><<<grid_for(total), 256, 0, s>>>(dy, w, dx, N, C, H, W, Ho, Wo, stride);
        else dw_dgrad_kernel<5><<<grid_for(total), 256, 0, s>>>(dy, w, dx, N, C, H, W, Ho, Wo, stride);
        YFV2_LAUNCH_CHECK();
    }
    if (dw) {
        YFV2_CUDA(cudaMemsetAsync(dw, 0, (size_t)C * ks * ks * sizeof(float), s));
        dim3 grid(C, N < 16 ? N : 16);
        if (ks == 3) dw_wgrad_kernel<3><<<grid, 256, 0, s>>>(x, dy, dw, N, C, H, W, Ho, Wo, stride);
        else dw_wgrad_kernel<5><<<grid, 256, 0, s>>>(x, dy, dw, N, C, H, W, Ho, Wo, stride);
        YFV2_LAUNCH_CHECK();
    }
    return YFV2_OK;
}

extern "C" YFV2_API int yfv2_op_stem_fwd(const float* x, const float* w, float* y, int N, int M, int H, int W, void* stream) {
    ARGCHK(x && w && y && H % 2 == 0 && W % 2 == 0, "stem_fwd: bad arguments");
    stem_fwd_kernel<<<grid_for((long long)N * M * (H / 2) * (W / 2)), 256, 0, (cudaStream_t)stream>>>(x, w, y, N, M, H, W);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}
extern "C" YFV2_API int yfv2_op_stem_wgrad(const float* x, const float* dy, float* dw, int N, int M, int H, int W, void* stream) {
    ARGCHK(x && dy && dw, "stem_wgrad: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    YFV2_CUDA(cudaMemsetAsync(dw, 0, (size_t)M * 27 * sizeof(float), s));
    stem_wgrad_kernel<<<dim3(M, 64), 256, 0, s>>>(x, dy, dw, N, M, H, W);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}

// scratch: 2*C doubles.  save_mean / save_invstd: [C] each.  running_* may be null (no update).
extern "C" YFV2_API int yfv2_op_bn_train_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var, float* y,
                                             float* save_mean, float* save_invstd, double* scratch, int N, int C, int HW, int relu, void* stream) {
    ARGCHK(x && gamma && beta && y && save_mean && save_invstd && scratch, "bn_train_fwd: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    YFV2_CUDA(cudaMemsetAsync(scratch, 0, (size_t)2 * C * sizeof(double), s));
    const long long per = (long long)N * HW;
    int slices = (int)((per + 256 * 16 - 1) / (256 * 16));
    slices = slices < 1 ? 1 : (slices > 64 ? 64 : slices);
    bn_stats_kernel<<<dim3(C, slices), 256, 0, s>>>(x, scratch, N, C, HW);
    YFV2_LAUNCH_CHECK();
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(scratch, save_mean, save_invstd, running_mean, running_var, C, (double)per, 0.1f);
    YFV2_LAUNCH_CHECK();
    const long long total = per * C;
    bn_apply_kernel<<<grid_for(total), 256, 0, s>>>(x, save_mean, save_invstd, gamma, beta, y, total, C, HW, relu);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}
extern "C" YFV2_API int yfv2_op_bn_train_bwd(const float* x, const float* y, const float* dy, const float* gamma, const float* save_mean,
                                             const float* save_invstd, float* dx, float* dgamma, float* dbeta, double* scratch, int N, int C, int HW,
                                             int relu, void* stream) {
    ARGCHK(x && y && dy && gamma && save_mean && save_invstd && dx && dgamma && dbeta && scratch, "bn_train_bwd: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    YFV2_CUDA(cudaMemsetAsync(scratch, 0, (size_t)2 * C * sizeof(double), s));
    const long long per = (long long)N * HW;
    int slices = (int)((per + 256 * 16 - 1) / (256 * 16));
    slices = slices < 1 ? 1 : (slices > 64 ? 64 : slices);
    bn_bwd_reduce_kernel<<<dim3(C, slices), 256, 0, s>>>(x, y, dy, save_mean, save_invstd, scratch, N, C, HW, relu);
    YFV2_LAUNCH_CHECK();
    const long long total = per * C;
    int grid = grid_for(total);
    if (grid < (C + 255) / 256) grid = (C + 255) / 256;
    bn_bwd_apply_kernel<<<grid, 256, 0, s>>>(x, y, dy, save_mean, save_invstd, gamma, scratch, dx, dgamma, dbeta, total, C, HW, (double)per, relu);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}

extern "C" YFV2_API int yfv2_op_maxpool_fwd(const float* x, float* y, int* idx, int planes, int H, int W, void* stream) {
    ARGCHK(x && y && idx, "maxpool_fwd: bad arguments");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    maxpool_fwd_kernel<<<grid_for((long long)planes * Ho * Wo), 256, 0, (cudaStream_t)stream>>>(x, y, idx, planes, H, W);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}
extern "C" YFV2_API int yfv2_op_maxpool_bwd(const float* dy, const int* idx, float* dx, int planes, int H, int W, void* stream) {
    ARGCHK(dy && idx && dx, "maxpool_bwd: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    YFV2_CUDA(cudaMemsetAsync(dx, 0, (size_t)planes * H * W * sizeof(float), s));
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    maxpool_bwd_kernel<<<grid_for((long long)planes * Ho * Wo), 256, 0, s>>>(dy, idx, dx, planes, H, W);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}
extern "C" YFV2_API int yfv2_op_upsample2_fwd(const float* x, float* y, int planes, int H, int W, void* stream) {
    ARGCHK(x && y, "upsample2_fwd: bad arguments");
    upsample2_fwd_kernel<<<grid_for((long long)planes * 4 * H * W), 256, 0, (cudaStream_t)stream>>>(x, y, planes, H, W);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}
extern "C" YFV2_API int yfv2_op_upsample2_bwd(const float* dy, float* dx, int planes, int H, int W, void* stream) {
    ARGCHK(dy && dx, "upsample2_bwd: bad arguments");
    upsample2_bwd_kernel<<<grid_for((long long)planes * H * W), 256, 0, (cudaStream_t)stream>>>(dy, dx, planes, H, W);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}
