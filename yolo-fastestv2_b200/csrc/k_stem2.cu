// K0 on the tensor core — stem: conv3x3 s2 p1 (3->24, no bias) + BN + ReLU + maxpool3x3 s2 p1 as a "strip walk".
// Reference: model/backbone/shufflenetv2.py:74-80,103-104; uint8 input fuses the `/255.0` of utils/utils.py:368.
//
// The FFMA2 stem (k_stem.cu) is bound by the fp32 pipe: 10.3 GFLOP per batch-256 launch is 138 us at 100 % of it (measured
// 407 us at 44 %), above the 124 us the 0.70-of-HBM target allows.  The first tensor-core stem (round 1, tc_stem_kernel) paid an
// im2col through shared memory, a CTA-wide producer/consumer chain and a shared-memory pooling pass.  This kernel has none of
// them:
//   * space-to-depth instead of im2col.  A TMEM lane is one 2x4 input block (rows 2y..2y+1, columns 4X..4X+3): it yields the
//     two conv positions (y, 2X) and (y, 2X+1), i.e. one column of the pooled map.  Its operand row is the 3 x 5 x 3 input
//     values those two positions touch (rows 2y-1..2y+1, columns 4X-1..4X+3, 3 channels: K = 45 -> 48), the weights are a
//     [48 x 48] matrix (two positions x 24 channels, zero where a tap does not reach) built in shared memory by the prologue.
//   * a WARP owns a strip of 31 pooled columns of one image band and walks down the conv rows: lane l <-> X = 31 s - 1 + l
//     (lane 0 duplicates the previous strip's last column, so the horizontal 3-max needs one shuffle and no exchange between
//     warps), the input row 2y+1 of a step is carried in registers as row 2(y+1)-1 of the next one, and the vertical 3-max is
//     a running maximum in registers: nothing is staged in, or pooled through, shared memory.
//   * four warps (four strips) share one M = 128 tcgen05.mma per step; the last warp to have stored its operand rows
//     (tcgen05.st) issues the MMAs.  uint8 pixels are exact in TF32, so the operand is not split (2 passes: A.Whi + A.Wlo,
//     1/255 folded into the weights); fp32 input uses the 3xTF32 scheme of tc.cuh (3 passes).
//   * BN scale is folded into the weights, the shift and the ReLU commute with max and are applied once per pooled value.
// Per step and lane: 6 loads (LDG.32 / LDG.128), 30 conversions, 3-6 tcgen05.st, 3 tcgen05.ld, 24 shuffles, ~100 max/ALU.
#include <type_traits>

#include "eng3.cuh"

namespace yfv2 {
namespace {
using namespace tc;
using namespace eng3;

struct Stem2Args {
    const void* x;
    Planes out;
    const float* wpack;      // STEM layout: Wt[27][24] (k = c*9+ky*3+kx) | scale[24] | shift[24]
    int N, H, W;
    int TRo, bands, nstrips; // pooled rows per band, bands per image, strips of 31 pooled columns per row
    int items;               // warp items = N * bands * nstrips
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                   "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

constexpr int kS2K = 48, kS2N = 48;        // operand row (45 used) / accumulator columns (2 positions x 24 channels)
constexpr int kStripW = 31;                // pooled columns per warp

// raw input of one lane for one input row: 4 pixels of each of the 3 channels.  The loads are PREDICATED (registers zeroed
// first) instead of select-after-load, so that nothing consumes them before the next step: a step's rows are requested one
// step ahead and fly during the MMAs and the epilogue.
template <bool U8> struct RawRow;
template <> struct RawRow<true> { uint32_t w[3]; };
template <> struct RawRow<false> { float4 v[3]; };

// ptr: this lane's address of (channel 0, the row) or nullptr-equivalent when !ok; cstride = H*W elements between channels
template <bool U8>
__device__ __forceinline__ void load_row(RawRow<U8>& r, const void* ptr, long long cstride, bool ok) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if constexpr (U8) {
            r.w[c] = 0u;
            if (ok) r.w[c] = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(ptr) + c * cstride));
        } else {
            r.v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) r.v[c] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ptr) + c * cstride));
        }
    }
}

__device__ __forceinline__ float byte_to_float(uint32_t w, int i) {       // exact: (2^23 + b) - 2^23
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | (uint32_t)i)) - 8388608.0f;
}

// the 15 operand values of one input row: [c][j], j = 0 <-> column 4X-1 (the left neighbour lane's last pixel), 1..4 <-> 4X..4X+3
template <bool U8>
__device__ __forceinline__ void row_values(const RawRow<U8>& r, float* v) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if constexpr (U8) {
            const uint32_t w = r.w[c];
            const uint32_t wl = __shfl_up_sync(0xffffffffu, w, 1);
            v[5 * c + 0] = byte_to_float(wl, 3);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[5 * c + 1 + j] = byte_to_float(w, j);
        } else {
            const float4 q = r.v[c];
            v[5 * c + 0] = __shfl_up_sync(0xffffffffu, q.w, 1);
            v[5 * c + 1] = q.x; v[5 * c + 2] = q.y; v[5 * c + 3] = q.z; v[5 * c + 4] = q.w;
        }
    }
}

template <bool U8, int G>
__global__ void __launch_bounds__(G * 128, 1)
stem2_kernel(const __grid_constant__ Stem2Args p) {
    pdl_trigger();
    constexpr int ACOLS = U8 ? kS2K : 2 * kS2K;             // operand columns (fp32 input: hi | lo)
    constexpr int COLS = ACOLS + kS2N;
    static_assert(G * COLS <= 512, "TMEM");
    __shared__ __align__(128) float sB[2 * kS2N * kS2K];    // Whi | Wlo, UMMA K-major no-swizzle tiles (tc.cuh)
    __shared__ float sShift[24];
    __shared__ __align__(8) BPipe pipes[G];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, grp = tid >> 7;

    {   // weights: B[k = (r, c, j)][n = (b, ch)] = W[ch][c][ky = r][kx = j - 2b] * bn_scale[ch]  (0 where the tap does not reach)
        const float* scale = p.wpack + 27 * 24;
        for (int i = tid; i < kS2N * kS2K; i += G * 128) {
            const int n = i / kS2K, k = i - n * kS2K;
            const int r = k / 15, rem = k - r * 15, c = rem / 5, j = rem - c * 5;
            const int b = n / 24, ch = n - b * 24, kx = j - 2 * b;
            float w = 0.f;
            if (k < 45 && kx >= 0 && kx <= 2) {
                w = __fmul_rn(__ldg(p.wpack + ((c * 3 + r) * 3 + kx) * 24 + ch), __ldg(scale + ch));
                if (U8) w = __fdiv_rn(w, 255.0f);
            }
            const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
            const int idx = tc_b_index(n, k, kS2K);
            sB[idx] = hi;
            sB[kS2N * kS2K + idx] = w - hi;
        }
        if (tid < 24) sShift[tid] = __ldg(scale + 24 + tid);
    }
    if (tid == 32) {
        for (int i = 0; i < G; ++i) {
            mbar_init(&pipes[i].empty[0], 1); mbar_init(&pipes[i].empty[1], 1); mbar_init(&pipes[i].dfull, 1);
            pipes[i].arrivals[0] = 0; pipes[i].arrivals[1] = 0;
        }
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    publish_async();                                        // the generic-proxy writes of sB before the tensor core reads them
    fence_before_sync();
    __syncthreads();
    fence_after_sync();

    BGrp g;
    g.tcol = tmem_slot + grp * COLS;
    g.tlane = g.tcol + ((uint32_t)(32 * (warp & 3)) << 16);
    g.pipe = &pipes[grp];
    g.chunk = 0; g.dparity = 0;
    g.gtid = tid & 127;
    const uint32_t b_hi = smem_u32(sB), b_lo = smem_u32(sB + kS2N * kS2K);
    const int H = p.H, W = p.W, HC = H / 2, HO = H / 4, WO = W / 4;
    const int gitems = (p.items + 3) >> 2;
    pdl_wait();                                             // the output planes may still be read by the previous forward's kernels

    constexpr int ESZ = U8 ? 1 : 4;                          // bytes per input element
    const long long cstride = (long long)H * W;             // elements between the channels of an image
    const float left_cap = -INFINITY;

    for (int gi = blockIdx.x * G + grp; gi < gitems; gi += gridDim.x * G) {
        const int wi = gi * 4 + (warp & 3);
        const bool wvalid = wi < p.items;
        const int strip = wi % p.nstrips, t1 = wi / p.nstrips;
        const int band = t1 % p.bands, n = wvalid ? t1 / p.bands : 0;
        const int X = kStripW * strip - 1 + lane;
        const bool xok = wvalid && X >= 0 && X < WO;
        const int oy0 = band * p.TRo;
        int y = 2 * oy0 - 1;                                // conv row of step 0
        // this lane's address of (channel 0, input row 2y-1, column 4X); walks down two input rows per step
        const char* rp = reinterpret_cast<const char*>(p.x) + ((long long)n * 3 * cstride + (long long)(2 * y - 1) * W + 4 * (xok ? X : 0)) * ESZ;
        const long long rstride = (long long)W * ESZ;
        float U[15];
        {
            RawRow<U8> r;
            load_row<U8>(r, rp, cstride, xok && 2 * y - 1 >= 0);
            row_values<U8>(r, U);
        }
        RawRow<U8> rm, rl;
        load_row<U8>(rm, rp + rstride, cstride, xok && y >= 0 && 2 * y < H);
        load_row<U8>(rl, rp + 2 * rstride, cstride, xok && y >= 0 && 2 * y + 1 < H);
        rp += 3 * rstride;                                  // -> row 2(y+1)
        float acc[24];
        float* orow = p.out.base + (long long)n * p.out.sN + p.out.org + (long long)oy0 * p.out.Ws + (xok ? X : 0);   // pooled row of the next emit
        const bool store_lane = xok && lane >= 1;
        const float lcap = X <= 0 ? left_cap : INFINITY;    // conv column -1 never wins a window
        int oy = oy0;

        // PH 0: first step of the band (acc = h)   1: middle row of a window (acc = max)   2: last row (emit, acc = h)
        auto step = [&](auto ph, bool prefetch) {
            constexpr int PH = decltype(ph)::value;
            // ---- operand row: [U | M | L | 0 0 0] -> TMEM ---------------------------------------------------------------
            {
                float a[48];
#pragma unroll
                for (int i = 0; i < 15; ++i) a[i] = U[i];
                row_values<U8>(rm, a + 15);
                row_values<U8>(rl, a + 30);
                a[45] = 0.f; a[46] = 0.f; a[47] = 0.f;
#pragma unroll
                for (int i = 0; i < 15; ++i) U[i] = a[30 + i];  // row 2y+1 is row 2(y+1)-1 of the next step
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    uint32_t hi[16];
                    if constexpr (U8) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) hi[i] = __float_as_uint(a[16 * q + i]);
                        tmem_st16(g.tlane + 16 * q, hi);
                    } else {
                        uint32_t lo[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            hi[i] = __float_as_uint(a[16 * q + i]) & 0xFFFFE000u;
                            lo[i] = __float_as_uint(a[16 * q + i] - __uint_as_float(hi[i]));
                        }
                        tmem_st16(g.tlane + 16 * q, hi);
                        tmem_st16(g.tlane + kS2K + 16 * q, lo);
                    }
                }
            }
            // ---- hand-off: the last of the group's four warps issues this step's MMAs -----------------------------------------
            wait_st();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                const uint32_t old = atom_inc_acq_rel(&g.pipe->arrivals[0]);
                if ((old & 3u) == 3u) {
                    fence_after_sync();
                    constexpr uint32_t idesc = make_idesc_tf32(128, kS2N);
                    constexpr uint32_t LBO = 128, SBO = (kS2K / 4) * 128;
                    const uint32_t d = g.tcol + ACOLS;
#pragma unroll
                    for (int s = 0; s < kS2K / 8; ++s) {
                        const uint64_t bh = make_b_desc(b_hi + s * 256, LBO, SBO);
                        const uint64_t bl = make_b_desc(b_lo + s * 256, LBO, SBO);
                        if constexpr (U8) {
                            mma_tf32_ts(d, g.tcol + 8 * s, bl, idesc, s > 0 ? 1u : 0u);      // small terms first
                            mma_tf32_ts(d, g.tcol + 8 * s, bh, idesc, 1u);
                        } else {
                            mma_tf32_ts(d, g.tcol + kS2K + 8 * s, bh, idesc, s > 0 ? 1u : 0u);
                            mma_tf32_ts(d, g.tcol + 8 * s, bl, idesc, 1u);
                            mma_tf32_ts(d, g.tcol + 8 * s, bh, idesc, 1u);
                        }
                    }
                    mma_commit(&g.pipe->dfull);
                }
            }
            __syncwarp();
            // ---- next step's input rows fly during the MMAs and the epilogue (rows past the image are not loaded) ------------
            {
                const bool more = prefetch && xok && 2 * y + 3 < H;     // row 2(y+1)+1 inside the image (then row 2(y+1) is too; y+1 >= 0 here)
                load_row<U8>(rm, rp, cstride, more);
                load_row<U8>(rl, rp + rstride, cstride, more);
                rp += 2 * rstride;
            }
            // ---- epilogue: horizontal 3-max (one shuffle), running vertical 3-max, shift + ReLU on the pooled value ------------
            mbar_wait(&g.pipe->dfull, g.dparity);
            g.dparity ^= 1u;
            fence_after_sync();
            float h[24];
            {
                float d[48];
#pragma unroll
                for (int n0 = 0; n0 < 48; n0 += 16) tmem_ld16v(g.tlane + ACOLS + n0, d + n0);
                wait_ld();
#pragma unroll
                for (int ch = 0; ch < 24; ++ch) {
                    const float left = fminf(__shfl_up_sync(0xffffffffu, d[24 + ch], 1), lcap);
                    h[ch] = fmaxf(fmaxf(left, d[ch]), d[24 + ch]);
                }
            }
            if (!(y >= 0 && y < HC)) {                      // conv rows outside the conv output never win a window (warp-uniform, rare)
#pragma unroll
                for (int ch = 0; ch < 24; ++ch) h[ch] = -INFINITY;
            }
            if constexpr (PH == 0) {
#pragma unroll
                for (int ch = 0; ch < 24; ++ch) acc[ch] = h[ch];
            } else if constexpr (PH == 1) {
#pragma unroll
                for (int ch = 0; ch < 24; ++ch) acc[ch] = fmaxf(acc[ch], h[ch]);
            } else {
                if (store_lane && oy < HO) {
                    float* op = orow;
#pragma unroll
                    for (int ch = 0; ch < 24; ++ch) {
                        *op = fmaxf(fmaxf(acc[ch], h[ch]) + sShift[ch], 0.f);
                        op += p.out.sC;
                    }
                }
#pragma unroll
                for (int ch = 0; ch < 24; ++ch) acc[ch] = h[ch];
                orow += p.out.Ws;
                ++oy;
            }
            ++y;
        };
        step(std::integral_constant<int, 0>{}, true);
#pragma unroll 1
        for (int i = 0; i < p.TRo; ++i) {
            step(std::integral_constant<int, 1>{}, true);
            step(std::integral_constant<int, 2>{}, i + 1 < p.TRo);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

}  // namespace

// false: geometry / alignment this kernel does not take (the caller falls back to the FFMA2 stem)
bool stem2_supported(const StemArgs& a) {
    if (a.H % 4 || a.W % 4 || a.H < 4 || a.W < 4) return false;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(a.x);
    return a.is_u8 ? (addr % 4 == 0) : (addr % 16 == 0);
}

template <bool U8, int G>
static int run_stem2(Stem2Args k, int N, int HO, cudaStream_t s) {
    const int slots = sm_count() * G;
    // bands: every extra band re-computes one conv row; more bands balance the grid.  Pick the band count with the smallest
    // (rounds x steps) product.
    long long best = -1;
    for (int bands = 1; bands <= HO; ++bands) {
        const int tro = (HO + bands - 1) / bands;
        if ((HO + tro - 1) / tro != bands) continue;
        const long long gitems = ((long long)N * bands * k.nstrips + 3) / 4;
        const long long cost = ((gitems + slots - 1) / slots) * (2 * tro + 1);
        if (best < 0 || cost < best) { best = cost; k.TRo = tro; k.bands = bands; }
    }
    k.items = N * k.bands * k.nstrips;
    const int gitems = (k.items + 3) / 4;
    const int ctas = (gitems + G - 1) / G;
    YFV2_CUDA(launch_k(stem2_kernel<U8, G>, ctas < sm_count() ? ctas : sm_count(), G * 128, 0, s, pdl_take(), k));
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}

int launch_stem2(const StemArgs& a, cudaStream_t s) {
    Stem2Args k{a.x, a.out, a.wpack, a.N, a.H, a.W, 0, 0, 0, 0};
    k.nstrips = (a.W / 4 + kStripW - 1) / kStripW;
    return a.is_u8 ? run_stem2<true, 4>(k, a.N, a.H / 4, s) : run_stem2<false, 3>(k, a.N, a.H / 4, s);
}

}  // namespace yfv2
