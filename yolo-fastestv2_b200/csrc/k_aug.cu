// Device-side `contrast_and_brightness` (reference utils/datasets.py:10-16, the only augmentation img_aug applies, :63-68):
//   dst = cv2.addWeighted(img, alpha, zeros, 1 - alpha, beta)  on uint8 images
// i.e. per element  saturate_cast<uint8>( round_half_even( fl32(fl32(x * alpha) + beta) ) )  (the zero image contributes
// exactly 0; OpenCV evaluates the weighted sum in fp32 and rounds with cvRound).  One (alpha, beta) pair per image, drawn by
// the caller (the reference draws them with random.uniform on the host).  Pure byte streaming: 16 bytes per thread and trip.
#include "common.cuh"

namespace yfv2 {
namespace {

__device__ __forceinline__ unsigned aug_byte(unsigned b, float alpha, float beta) {
    const float t = __fadd_rn(__fmul_rn((float)b, alpha), beta);
    const float r = fminf(fmaxf(rintf(t), 0.f), 255.f);      // rintf: round half to even (cvRound)
    return (unsigned)r;
}
__device__ __forceinline__ unsigned aug_word(unsigned w, float alpha, float beta) {
    return aug_byte(w & 0xffu, alpha, beta) | (aug_byte((w >> 8) & 0xffu, alpha, beta) << 8) |
           (aug_byte((w >> 16) & 0xffu, alpha, beta) << 16) | (aug_byte(w >> 24, alpha, beta) << 24);
}

// img / out: [N][per_image] bytes; alpha / beta: [N]
__global__ void __launch_bounds__(256)
contrast_brightness_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, const float* __restrict__ alpha,
                           const float* __restrict__ beta, long long per_image, int vec_ok) {
    const int n = blockIdx.y;
    const float a = __ldg(alpha + n), b = __ldg(beta + n);
    const uint8_t* src = img + (long long)n * per_image;
    uint8_t* dst = out + (long long)n * per_image;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec_ok) {
        const long long nv = per_image >> 4;
        for (long long v = i; v < nv; v += stride) {
            uint4 q = __ldg(reinterpret_cast<const uint4*>(src) + v);
            q.x = aug_word(q.x, a, b); q.y = aug_word(q.y, a, b); q.z = aug_word(q.z, a, b); q.w = aug_word(q.w, a, b);
            reinterpret_cast<uint4*>(dst)[v] = q;
        }
        for (long long e = (nv << 4) + i; e < per_image; e += stride) dst[e] = (uint8_t)aug_byte(src[e], a, b);
    } else {
        for (long long e = i; e < per_image; e += stride) dst[e] = (uint8_t)aug_byte(src[e], a, b);
    }
}

}  // namespace
}  // namespace yfv2

extern "C" int yfv2_aug_contrast_brightness(const uint8_t* img, uint8_t* out, const float* alpha, const float* beta, int N,
                                            long long bytes_per_image, void* stream) {
    using namespace yfv2;
    if (!img || !out || !alpha || !beta || N <= 0 || bytes_per_image <= 0) { set_error("aug_contrast_brightness: bad argument"); return YFV2_EINVAL; }
    const int vec_ok = (reinterpret_cast<uintptr_t>(img) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0) && (bytes_per_image % 16 == 0);
    long long blocks = (bytes_per_image / 16 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
    contrast_brightness_kernel<<<dim3((unsigned)blocks, (unsigned)N), 256, 0, (cudaStream_t)stream>>>(img, out, alpha, beta, bytes_per_image, vec_ok);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}
