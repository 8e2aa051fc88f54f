// Bring-up / regression kernel for the tcgen05 pointwise engine (tc.cuh): out[n][p] = sum_k w[n][k] * x[k][p]
// for one 128-pixel tile per CTA.  Exposed as yfv2_debug_pw_tc so a GPU test can pin descriptor encodings, TMEM
// addressing and the 3xTF32 accuracy against an fp32 reference before the fused kernels rely on them.
#include "common.cuh"
#include "tc.cuh"

namespace yfv2 {
namespace {

__global__ void pack_tc_kernel(const float* __restrict__ w, int N, int K, int NP, int KP, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NP * KP) return;
    const int n = i / KP, k = i - n * KP;
    const float v = (n < N && k < K) ? w[n * K + k] : 0.f;
    const uint32_t hi = tc::tf32_rna(v);
    const uint32_t lo = tc::tf32_rna(v - __uint_as_float(hi));
    const int o = tc::tc_b_index(n, k, KP);
    dst[o] = __uint_as_float(hi);
    dst[NP * KP + o] = __uint_as_float(lo);
}

template <int K, int N>
__global__ void __launch_bounds__(128)
tc_pw_test_kernel(const float* __restrict__ x, const float* __restrict__ pack, float* __restrict__ out, int P) {
    constexpr int KP = tc::tc_round(K, 8), NP = tc::tc_round(N, 16);
    constexpr int NEED = 2 * KP + NP;
    constexpr uint32_t NCOLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
    extern __shared__ __align__(128) float smem[];
    float* sB = smem;                                   // Bhi | Blo
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 2 * NP * KP; i += 128) sB[i] = __ldg(pack + i);
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, NCOLS);
    if (tid == 0) { tc::mbar_init(&mbar, 1); tc::fence_mbar_init(); }
    // make the generic-proxy writes of B visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(32 * (warp & 3)) << 16);

    const int p = blockIdx.x * 128 + tid;
#pragma unroll
    for (int k0 = 0; k0 < KP; k0 += 8) {
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = (p < P && k0 + j < K) ? __ldg(x + (size_t)(k0 + j) * P + p) : 0.f;
        tc::store_a8(lane_base + k0, KP, a);
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after_sync();
        tc::issue_pw<KP, NP, KP>(tbase + 2 * KP, tbase, tbase + KP, tc::smem_u32(sB), tc::smem_u32(sB + NP * KP), 0, false);
        tc::mma_commit(&mbar);
    }
    tc::mbar_wait(&mbar, 0);
    tc::fence_after_sync();
#pragma unroll
    for (int n0 = 0; n0 < NP; n0 += 8) {
        float d[8];
        tc::tmem_ld8(lane_base + 2 * KP + n0, d);
        tc::wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (p < P && n0 + j < N) out[(size_t)(n0 + j) * P + p] = d[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, NCOLS);
}

template <int K, int N>
int run_test(const float* x, const float* w, float* out, float* pack_ws, int P, cudaStream_t s) {
    constexpr int KP = tc::tc_round(K, 8), NP = tc::tc_round(N, 16);
    pack_tc_kernel<<<(NP * KP + 255) / 256, 256, 0, s>>>(w, N, K, NP, KP, pack_ws);
    YFV2_LAUNCH_CHECK();
    const size_t bytes = (size_t)2 * NP * KP * sizeof(float);
    auto kern = tc_pw_test_kernel<K, N>;
    YFV2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    kern<<<(P + 127) / 128, 128, bytes, s>>>(x, pack_ws, out, P);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}
}  // namespace
}  // namespace yfv2

using namespace yfv2;

extern "C" int yfv2_debug_pw_tc(const float* x, const float* w, float* out, float* pack_ws, int K, int N, int P, void* stream) {
    if (!x || !w || !out || !pack_ws || P <= 0) { set_error("debug_pw_tc: bad argument"); return YFV2_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    if (K == 24 && N == 24) return run_test<24, 24>(x, w, out, pack_ws, P, s);
    if (K == 48 && N == 48) return run_test<48, 48>(x, w, out, pack_ws, P, s);
    if (K == 96 && N == 96) return run_test<96, 96>(x, w, out, pack_ws, P, s);
    if (K == 72 && N == 72) return run_test<72, 72>(x, w, out, pack_ws, P, s);
    if (K == 72 && N == 83) return run_test<72, 83>(x, w, out, pack_ws, P, s);
    set_error("debug_pw_tc: unsupported K=%d N=%d", K, N);
    return YFV2_EUNSUPPORTED;
}
