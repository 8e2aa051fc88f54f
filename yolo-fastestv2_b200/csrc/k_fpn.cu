// K3 — LightFPN reducers (reference model/fpn.py:35-43,52,57-59):
//   S3 = ReLU(BN(conv1x1 192->72 (C3)))
//   S2 = ReLU(BN(conv1x1 288->72 (cat(nearest_up2(C3), C2))))
// The upsample and the concat are never materialised: a thread owning output pixel (y,x) reads C3 at
// (y>>1, x>>1) for k < 192 and C2 at (y,x) for the rest ("gather on load"), straight from the channel
// planes the backbone left behind.  One thread = one output pixel x 72 output channels (register
// accumulators); the weight rows are float4 broadcasts from shared memory.
#include "common.cuh"

namespace yfv2 {
namespace {
constexpr int NT = 256;
constexpr int NOUT = 72;

// KA planes read through (y>>SHA, x>>SHA) from A, then KB planes read at (y,x) from B.
template <int KA, int KB, int SHA>
__global__ void __launch_bounds__(NT)
fpn_pw_kernel(Planes A, ChanTab ta, Planes B, ChanTab tb, Planes out, const float* __restrict__ wpack, int total) {
    extern __shared__ __align__(16) float smem[];
    constexpr int K = KA + KB;
    copy_to_smem(smem, wpack, pw_pack_floats(K, NOUT));
    __syncthreads();
    const float* scale = smem + K * NOUT;
    const float* shift = scale + NOUT;
    const int HW = out.H * out.W, W = out.W;
    const int gid = blockIdx.x * NT + threadIdx.x;
    if (gid >= total) return;
    const int n = gid / HW, p = gid - n * HW;
    const int y = p / W, x = p - y * W;
    const long long offA = (long long)(y >> SHA) * A.W + (x >> SHA);
    const float* baseA = A.base + (long long)n * A.sN + offA;
    const float* baseB = KB ? B.base + (long long)n * B.sN + p : nullptr;

    float acc[1][NOUT];
#pragma unroll
    for (int j = 0; j < NOUT; ++j) acc[0][j] = 0.f;
    constexpr int U = 8;
    static_assert(KA % U == 0 && KB % U == 0, "unroll");
    for (int k0 = 0; k0 < KA; k0 += U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(baseA + (long long)ta.c[k0 + u] * A.sC);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float xv[1] = {v[u]};
            fma_row<NOUT, 1>(smem + (k0 + u) * NOUT, xv, acc);
        }
    }
    for (int k0 = 0; k0 < KB; k0 += U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(baseB + (long long)tb.c[k0 + u] * B.sC);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float xv[1] = {v[u]};
            fma_row<NOUT, 1>(smem + (KA + k0 + u) * NOUT, xv, acc);
        }
    }
    float* o = out.base + (long long)n * out.sN + p;
#pragma unroll
    for (int j = 0; j < NOUT; ++j) o[(long long)j * out.sC] = fmaxf(fmaf(acc[0][j], scale[j], shift[j]), 0.f);
}
}  // namespace

int launch_fpn(const FpnArgs& a, int which, cudaStream_t s) {
    if (which == 0) {
        auto kern = fpn_pw_kernel<192, 0, 0>;
        const size_t bytes = pw_pack_floats(192, NOUT) * sizeof(float);
        YFV2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        const int total = a.N * a.s3.H * a.s3.W;
        kern<<<(total + NT - 1) / NT, NT, bytes, s>>>(a.c3, a.t3, a.c3, a.t3, a.s3, a.w3, total);
        YFV2_LAUNCH_CHECK();
    } else {
        auto kern = fpn_pw_kernel<192, 96, 1>;
        const size_t bytes = pw_pack_floats(288, NOUT) * sizeof(float);
        YFV2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        const int total = a.N * a.s2.H * a.s2.W;
        kern<<<(total + NT - 1) / NT, NT, bytes, s>>>(a.c3, a.t3, a.c2, a.t2, a.s2, a.w2, total);
        YFV2_LAUNCH_CHECK();
    }
    return YFV2_OK;
}

}  // namespace yfv2
