// Host side of libyfv2.so: plan construction (shape bookkeeping, channel-plane tables, packed-weight and
// workspace layouts), weight packing, and the C ABI declared in include/yfv2.h.
//
// Mirrors the wiring of the reference model (model/detector.py:8-47, model/fpn.py:31-64,
// model/backbone/shufflenetv2.py:65-109) without any of its module objects: the plan is a flat list of
// fused-kernel launches over plane pools.
#include <stdarg.h>
#include <string.h>

#include <new>
#include <vector>

#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace yfv2 {

bool blk_s1_chainable(int K, int H, int W);      // k_blk.cu
bool tail_s1_supported(int K, int H, int W);     // k_tail.cu

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static thread_local bool g_pdl_next = false;
static thread_local bool g_pdl_call = true;      // false inside yfv2_detect_u8_host: with copies and a second stream in flight the
                                                 // early-resident dependents cost more than the hidden prologues save (measured)
static const bool g_pdl_off = getenv("YFV2_NO_PDL") != nullptr;
bool pdl_take() { const bool r = g_pdl_next && g_pdl_call && !g_pdl_off; g_pdl_next = true; return r; }
void pdl_reset() { g_pdl_next = false; }
bool pdl_allowed() { return g_pdl_call && !g_pdl_off; }

int sm_count() {
    static thread_local int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// pack sizes of the CUDA-core layouts (the depthwise packs inside them feed the tensor-core kernels too)
size_t shuffle_pack_floats(int K, int stride) {
    const size_t pwn = pw_pack_floats(K, K), dwn = dw3_pack_floats(K);
    return stride == 1 ? 2 * pwn + dwn : 3 * pwn + 2 * dwn;
}
size_t head_pack_floats() { return 2 * ((size_t)dw5_pack_floats(72) + pw_pack_floats(72, 72)); }

namespace {

constexpr int kStageRepeats[3] = {4, 8, 4};          // shufflenetv2.py:69
constexpr int kStageWidth[3] = {24, 48, 96};         // branch width = out_channels / 2 (detector.py:11)
constexpr int kNumBlocks = 16;
constexpr int kFpnDepth = 72;                        // detector.py:10

// ---- pack kernels ---------------------------------------------------------------------------------------
// w: [Nout][K] row-major conv weight.  Writes Wt[k][n_off+n], scale/shift[n_off+n] with row length Np.
// BN folded as PyTorch's eval path does: alpha = invstd*gamma, beta' = beta - mean*alpha.
__global__ void pack_pw_kernel(const float* __restrict__ w, int Nout, int K, const float* gamma, const float* beta,
                               const float* mean, const float* var, const float* bias, float* Wt, int Np, int n_off,
                               float* scale, float* shift) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Nout * K) {
        const int n = i / K, k = i - n * K;
        Wt[(size_t)k * Np + n_off + n] = w[i];
    }
    if (i < Nout) {
        float sc = 1.0f, sh = 0.0f;
        if (gamma) {
            const float invstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[i], kBnEps)));
            sc = __fmul_rn(invstd, gamma[i]);
            sh = __fsub_rn(beta[i], __fmul_rn(mean[i], sc));
        } else if (bias) {
            sh = bias[i];
        }
        scale[n_off + i] = sc;
        shift[n_off + i] = sh;
    }
}

// w: [C][KK] depthwise weight -> per channel [KK taps][scale][shift][pad] with row length R
__global__ void pack_dw_kernel(const float* __restrict__ w, int Cn, int KK, int R, const float* gamma, const float* beta,
                               const float* mean, const float* var, float* dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Cn * KK) {
        const int c = i / KK, t = i - c * KK;
        dst[(size_t)c * R + t] = w[i];
    }
    if (i < Cn) {
        const float invstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[i], kBnEps)));
        const float sc = __fmul_rn(invstd, gamma[i]);
        dst[(size_t)i * R + KK] = sc;
        dst[(size_t)i * R + KK + 1] = __fsub_rn(beta[i], __fmul_rn(mean[i], sc));
    }
}

// Tensor-core pack of a pointwise layer (tc.cuh): tf32 hi/lo split, UMMA K-major no-swizzle core-matrix tiling.
__global__ void pack_tc_kernel(const float* __restrict__ w, int Nout, int K, int NP, int KP, int n_off, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nout * KP) return;
    const int n = i / KP, k = i - n * KP;
    const float v = (k < K) ? w[n * K + k] : 0.f;
    const uint32_t hi = tc::tf32_rna(v);
    const uint32_t lo = tc::tf32_rna(v - __uint_as_float(hi));
    const int o = tc::tc_b_index(n_off + n, k, KP);
    dst[o] = __uint_as_float(hi);
    dst[NP * KP + o] = __uint_as_float(lo);
}

// Heads, second half (reference fpn.py DWConvblock tail + detector.py:17-19,35-41): pw(72->72), its BN and the shared output
// conv are consecutive affine maps with no activation between them, so they are ONE matrix at inference time:
//   F = Wout . diag(sc) . Wpw,   f = Wout . sh + bias          (products accumulated in fp64, rounded once to fp32)
// rows [n_off, n_off+Mo) of the folded [96 x K] matrix Fw and of the bias Fb.
__global__ void fold_head_kernel(const float* __restrict__ wout, const float* __restrict__ bout, int Mo, const float* __restrict__ wpw,
                                 const float* gamma, const float* beta, const float* mean, const float* var, int K, int n_off,
                                 float* __restrict__ Fw, float* __restrict__ Fb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Mo * K) {
        const int m = i / K, k = i - m * K;
        double acc = 0.0;
        for (int j = 0; j < K; ++j) {
            const float invstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[j], kBnEps)));
            const float sc = __fmul_rn(invstd, gamma[j]);
            acc += (double)wout[m * K + j] * (double)sc * (double)wpw[j * K + k];
        }
        Fw[(size_t)(n_off + m) * K + k] = (float)acc;
    }
    if (i < Mo) {
        double acc = (double)bout[i];
        for (int j = 0; j < K; ++j) {
            const float invstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[j], kBnEps)));
            const float sc = __fmul_rn(invstd, gamma[j]);
            const float sh = __fsub_rn(beta[j], __fmul_rn(mean[j], sc));
            acc += (double)wout[i * K + j] * (double)sh;
        }
        Fb[n_off + i] = (float)acc;
    }
}
__global__ void fold_affine_kernel(const float* __restrict__ Fb, int NP, float* __restrict__ aff) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < NP) { aff[i] = 1.0f; aff[NP + i] = Fb[i]; }
}

// dense NCHW copy of a logical tensor (tests / debugging only)
__global__ void gather_kernel(Planes P, ChanTab tab, int Cn, float* __restrict__ out, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int HW = P.H * P.W;
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % Cn);
    const int n = (int)(i / ((long long)HW * Cn));
    const int y = p / P.W, x = p - y * P.W;
    out[i] = plane_ptr(P, n, tab.c[c])[P.org + y * P.Ws + x];
}

struct Cursor {           // walks parameters / BN layers in state_dict order
    const float* const* params;
    const float* const* bn;
    int p = 0, b = 0;
};

}  // namespace
}  // namespace yfv2

using namespace yfv2;

struct yfv2_plan {
    int device, N, H, W, A, C, training;
    int h[4], w[4];                    // strides 4, 8, 16, 32
    size_t plane[4];                   // pool plane stride (floats) per resolution, zero frame of 1 included
    size_t plane2[4];                  // same with a frame of 2 (inputs of the 5x5 depthwise heads)
    int ws1[4], ws2[4];                // row strides for the two frame widths
    const void* ws_zeroed;             // workspace whose frames have been zeroed
    int pool_planes[4];                // 24, 72, 144, 288
    size_t off_pool[4];
    size_t off_s2, off_s3, off_t[4];   // t: cls2, reg2, cls3, reg3 mid-block scratch
    size_t ws_floats;
    size_t pk_stem, pk_block[kNumBlocks], pk_fpn3, pk_fpn2, pk_head[4], pk_out_reg, pk_out_oc, pk_floats;
    int blk_K[kNumBlocks], blk_stride[kNumBlocks], blk_res[kNumBlocks];   // res = index of OUTPUT resolution
    ChanTab tin[kNumBlocks], tout[kNumBlocks];
    ChanTab c2, c3;
    ChanTab logical[kNumBlocks];       // logical channel order of each block's output (debug gather)
    int launches;
    // tensor-core engine
    size_t tk_blk[kNumBlocks][3];      // tc packs per block: [0]=pw1 [1]=pw2 [2]=proj pw (stride-2 only)
    size_t tk_stem, tk_fpn3, tk_fpn2, tk_head[4][2], tk_out_oc, tk_out_reg;
    size_t tk_fold[4], pk_foldw[4];    // heads' second pointwise + BN + output conv folded into one matrix (tc pack / fp32 scratch)
    ChanTab t96;                       // scratch plane ids for the K=96 blocks' pw1 output
    int n_stages;
    struct Stage { int kind, a, b; char name[24]; int group; } stages[64];   // group: first stage of the launch this stage shares
};

namespace {

Planes pool_planes(const yfv2_plan* p, float* ws, int r) {
    Planes P;
    P.base = ws + p->off_pool[r];
    P.sC = (long long)p->plane[r];
    P.sN = (long long)p->plane[r] * p->pool_planes[r];
    P.H = p->h[r]; P.W = p->w[r];
    P.Ws = p->ws1[r]; P.pad = 1; P.org = P.Ws + 1;
    return P;
}
Planes flat_planes(const yfv2_plan* p, float* ws, size_t off, int r) {
    Planes P;
    P.base = ws + off;
    P.sC = (long long)p->plane2[r];
    P.sN = (long long)p->plane2[r] * kFpnDepth;
    P.H = p->h[r]; P.W = p->w[r];
    P.Ws = p->ws2[r]; P.pad = 2; P.org = 2 * P.Ws + 2;
    return P;
}

void build_tables(yfv2_plan* p) {
    std::vector<int> L(24);
    for (int i = 0; i < 24; ++i) L[i] = i;
    int bi = 0;
    for (int st = 0; st < 3; ++st) {
        const int K = kStageWidth[st];
        std::vector<int> freep;
        for (int rep = 0; rep < kStageRepeats[st]; ++rep, ++bi) {
            p->blk_K[bi] = K;
            p->blk_res[bi] = st + 1;
            if (rep == 0) {
                // stride 2: read every logical channel of the previous stage, write 2K fresh planes 0..2K-1
                p->blk_stride[bi] = 2;
                for (int i = 0; i < K; ++i) p->tin[bi].c[i] = (unsigned short)L[i];
                L.resize(2 * K);
                for (int i = 0; i < 2 * K; ++i) { L[i] = i; p->tout[bi].c[i] = (unsigned short)i; }
                freep.clear();
                for (int i = 2 * K; i < 3 * K; ++i) freep.push_back(i);
            } else {
                // stride 1: channel_shuffle (shufflenetv2.py:57-63): even logical channels pass through,
                // odd ones feed branch_main; output = cat(pass, main)
                p->blk_stride[bi] = 1;
                std::vector<int> pass, mainin;
                for (int i = 0; i < 2 * K; ++i) (i % 2 ? mainin : pass).push_back(L[i]);
                for (int i = 0; i < K; ++i) {
                    p->tin[bi].c[i] = (unsigned short)mainin[i];
                    p->tout[bi].c[i] = (unsigned short)freep[i];
                }
                L = pass;
                L.insert(L.end(), freep.begin(), freep.end());
                freep = mainin;
            }
            for (size_t i = 0; i < L.size(); ++i) p->logical[bi].c[i] = (unsigned short)L[i];
        }
        if (st == 1) for (int i = 0; i < 96; ++i) p->c2.c[i] = (unsigned short)L[i];
        if (st == 2) for (int i = 0; i < 192; ++i) p->c3.c[i] = (unsigned short)L[i];
    }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" int yfv2_abi_version(void) { return YFV2_ABI_VERSION; }
extern "C" const char* yfv2_last_error(void) { return g_err; }

extern "C" int yfv2_plan_create(yfv2_plan** out, int device, int N, int H, int W, int A, int C, int training) {
    if (!out || N <= 0 || H <= 0 || W <= 0 || H % 32 || W % 32 || A <= 0 || A > 8 || C <= 0 || C > 256) {
        set_error("plan_create: bad arguments (N=%d H=%d W=%d A=%d C=%d); H and W must be multiples of 32, A<=8, C<=256",
                  N, H, W, A, C);
        return YFV2_EINVAL;
    }
    if (training) {
        set_error("plan_create: training plans (batch-statistics BN + backward) are not implemented in this build");
        return YFV2_EUNSUPPORTED;
    }
    yfv2_plan* p = new (std::nothrow) yfv2_plan();
    if (!p) { set_error("plan_create: out of host memory"); return YFV2_ENOMEM; }
    memset(p, 0, sizeof(*p));
    p->device = device; p->N = N; p->H = H; p->W = W; p->A = A; p->C = C; p->training = training;
    // + scratch planes for the pw1 output of the blocks that run pointwise-1 as its own launch: K=96 (stage 4) and, with
    // YFV2_S2_48_SPLIT=1, the K=48 stride-2 block (stage3.0; scratch in the stride-8 pool)
    static const bool s2_48_split = getenv("YFV2_S2_48_SPLIT") != nullptr;
    const int pools[4] = {24, 72 + (s2_48_split ? 48 : 0), 144 + 96, 288 + 96};
    size_t off = 0;
    for (int r = 0; r < 4; ++r) {
        p->h[r] = H >> (r + 2); p->w[r] = W >> (r + 2);
        p->ws1[r] = (int)align_up(p->w[r] + 2, 4); p->ws2[r] = (int)align_up(p->w[r] + 4, 4);
        p->plane[r] = (size_t)(p->h[r] + 2) * p->ws1[r];
        p->plane2[r] = (size_t)(p->h[r] + 4) * p->ws2[r];
        p->pool_planes[r] = pools[r];
        p->off_pool[r] = off;
        off += align_up(p->plane[r] * pools[r] * (size_t)N, 64);
    }
    p->off_s2 = off; off += align_up(p->plane2[2] * kFpnDepth * (size_t)N, 64);
    p->off_s3 = off; off += align_up(p->plane2[3] * kFpnDepth * (size_t)N, 64);
    for (int i = 0; i < 4; ++i) {
        p->off_t[i] = off;
        off += align_up(p->plane2[i < 2 ? 2 : 3] * kFpnDepth * (size_t)N, 64);
    }
    p->ws_floats = off;
    build_tables(p);
    size_t pk = 0;
    p->pk_stem = pk; pk += align_up(kStemPackFloats, 4);
    for (int b = 0; b < kNumBlocks; ++b) { p->pk_block[b] = pk; pk += align_up(shuffle_pack_floats(p->blk_K[b], p->blk_stride[b]), 4); }
    p->pk_fpn3 = pk; pk += pw_pack_floats(192, kFpnDepth);
    p->pk_fpn2 = pk; pk += pw_pack_floats(288, kFpnDepth);
    for (int i = 0; i < 4; ++i) { p->pk_head[i] = pk; pk += align_up(head_pack_floats(), 4); }
    p->pk_out_reg = pk; pk += pw_pack_floats(kFpnDepth, 4 * A);
    p->pk_out_oc = pk; pk += pw_pack_floats(kFpnDepth, A + C);
    // tensor-core packs
    for (int b = 0; b < kNumBlocks; ++b) {
        const int K = p->blk_K[b], NP = tc::tc_round(K, 16);
        const int n = p->blk_stride[b] == 2 ? 3 : 2;
        for (int i = 0; i < n; ++i) { p->tk_blk[b][i] = pk; pk += align_up(2 * (size_t)NP * K + 2 * NP, 4); }
    }
    p->tk_stem = pk; pk += align_up(2 * 32 * 32 + 64, 4);
    p->tk_fpn3 = pk; pk += align_up(2 * 80 * 192 + 160, 4);
    p->tk_fpn2 = pk; pk += align_up(2 * 80 * 288 + 160, 4);
    for (int i = 0; i < 4; ++i) for (int h = 0; h < 2; ++h) { p->tk_head[i][h] = pk; pk += align_up(2 * 80 * 72 + 160, 4); }
    p->tk_out_oc = pk; pk += align_up(2 * 96 * 80 + 192, 4);
    p->tk_out_reg = pk; pk += align_up(2 * 96 * 80 + 192, 4);
    for (int i = 0; i < 4; ++i) { p->tk_fold[i] = pk; pk += align_up(2 * 96 * 72 + 192, 4); }
    for (int i = 0; i < 4; ++i) { p->pk_foldw[i] = pk; pk += align_up(96 * 72 + 96, 4); }
    p->pk_floats = pk;
    for (int i = 0; i < 96; ++i) p->t96.c[i] = 0;   // filled per use (pool-specific offset)

    if (A + C > 96 || 4 * A > 96) {
        set_error("plan_create: anchors+classes = %d exceeds the 96-wide output-conv tile of this build", A + C);
        delete p;
        return YFV2_EUNSUPPORTED;
    }
    auto add = [&](int kind, int a, int b, const char* fmt, int x, int y) {
        yfv2_plan::Stage& st = p->stages[p->n_stages++];
        st.kind = kind; st.a = a; st.b = b;
        snprintf(st.name, sizeof(st.name), fmt, x, y);
    };
    add(0, 0, 0, "stem", 0, 0);
    for (int b = 0, st = 0, rep = 0; b < kNumBlocks; ++b) {
        // experiment kept behind YFV2_S2_48_SPLIT=1: the K=48 stride-2 block as pw1 to scratch planes + channel-streamed dw->pw
        // branches (two launches, like stage4.0).  Measured 168.8 us against 139.5 us for the fused banded kernel (the extra
        // round trip of the pointwise-1 output through L2/HBM costs more than the idle warpgroups of the fused phases).
        if (p->blk_K[b] == 48 && p->blk_stride[b] == 2 && s2_48_split && p->h[p->blk_res[b]] * p->w[p->blk_res[b]] <= 512) {
            add(12, b, 0, "stage%d.%d/pw1", st + 2, rep); add(13, b, 0, "stage%d.%d/dwpw", st + 2, rep);
        } else if (p->blk_K[b] < 96) add(p->blk_stride[b] == 2 ? 11 : 10, b, 0, "stage%d.%d", st + 2, rep);
        else if (p->blk_stride[b] == 1 && tail_s1_supported(96, p->h[3], p->w[3])) add(16, b, 0, "stage%d.%d", st + 2, rep);   // small-map chain
        else { add(12, b, 0, "stage%d.%d/pw1", st + 2, rep); add(13, b, 0, "stage%d.%d/dwpw", st + 2, rep); }
        if (++rep == kStageRepeats[st]) { rep = 0; ++st; }
    }
    add(14, 1, 0, "fpn.S3", 0, 0); add(14, 2, 0, "fpn.S2", 0, 0);
    for (int lv = 0; lv < 2; ++lv) { add(15, lv, 0, "heads%d.a", lv + 2, 0); add(15, lv, 1, "heads%d.b", lv + 2, 0); }
    // launch groups: consecutive stride-1 blocks of a stage run as one chained launch when an image fits (k_blk.cu)
    p->launches = 0;
    for (int i = 0; i < p->n_stages; ++i) {
        yfv2_plan::Stage& st = p->stages[i];
        st.group = i;
        if (i > 0 && st.kind == 10 && p->stages[i - 1].kind == 10 && p->blk_K[st.a] == p->blk_K[p->stages[i - 1].a] &&
            i - p->stages[i - 1].group < 7 && blk_s1_chainable(p->blk_K[st.a], p->h[p->blk_res[st.a]], p->w[p->blk_res[st.a]]))
            st.group = p->stages[i - 1].group;
        if (i > 0 && st.kind == 16 && p->stages[i - 1].kind == 16 && i - p->stages[i - 1].group < 4) st.group = p->stages[i - 1].group;
        if (st.group == i) ++p->launches;
    }
    *out = p;
    return YFV2_OK;
}

extern "C" int yfv2_plan_destroy(yfv2_plan* p) {
    delete p;
    return YFV2_OK;
}

extern "C" int yfv2_plan_workspace_bytes(const yfv2_plan* p, size_t* bytes) {
    if (!p || !bytes) { set_error("workspace_bytes: null argument"); return YFV2_EINVAL; }
    *bytes = p->ws_floats * sizeof(float);
    return YFV2_OK;
}

extern "C" int yfv2_plan_packed_bytes(const yfv2_plan* p, size_t* bytes) {
    if (!p || !bytes) { set_error("packed_bytes: null argument"); return YFV2_EINVAL; }
    *bytes = p->pk_floats * sizeof(float);
    return YFV2_OK;
}

extern "C" const char* yfv2_plan_stage_name(const yfv2_plan* p, int i) {
    if (!p || i < 0 || i >= p->n_stages) return nullptr;
    return p->stages[i].name;
}

extern "C" int yfv2_plan_invalidate_workspace(yfv2_plan* p) {
    if (!p) { set_error("invalidate_workspace: null plan"); return YFV2_EINVAL; }
    p->ws_zeroed = nullptr;          // the next forward re-zeroes the frames around every activation plane
    return YFV2_OK;
}

extern "C" int yfv2_plan_stage_group(const yfv2_plan* p, int i) {
    if (!p || i < 0 || i >= p->n_stages) return -1;
    return p->stages[i].group;
}

extern "C" int yfv2_plan_forward_launches(const yfv2_plan* p, int* n) {
    if (!p || !n) { set_error("forward_launches: null argument"); return YFV2_EINVAL; }
    *n = p->launches;
    return YFV2_OK;
}

namespace {

struct TcDst { float* dst; int NP, KP; };

int pack_pw(Cursor& cur, bool bn, bool bias, int Nout, int K, float* pack, int Np, int n_off, cudaStream_t s, TcDst tcd = {nullptr, 0, 0}) {
    const float* w = cur.params[cur.p];
    const float *g = nullptr, *b = nullptr, *m = nullptr, *v = nullptr, *bi = nullptr;
    if (bn) { g = cur.params[cur.p + 1]; b = cur.params[cur.p + 2]; m = cur.bn[2 * cur.b]; v = cur.bn[2 * cur.b + 1]; cur.p += 3; cur.b += 1; }
    else if (bias) { bi = cur.params[cur.p + 1]; cur.p += 2; }
    else cur.p += 1;
    if (!w || (bn && (!g || !b || !m || !v)) || (bias && !bi)) { set_error("pack_weights: null tensor pointer near param %d", cur.p); return YFV2_EINVAL; }
    const int total = Nout * K;
    pack_pw_kernel<<<(total + 255) / 256, 256, 0, s>>>(w, Nout, K, g, b, m, v, bi, pack, Np, n_off, pack + (size_t)K * Np,
                                                        pack + (size_t)K * Np + Np);
    YFV2_LAUNCH_CHECK();
    if (tcd.dst) {
        // same layer for the tcgen05 engine: Bhi | Blo tiles, then scale[NP] | shift[NP] (zero padded by the memset)
        pack_tc_kernel<<<(Nout * tcd.KP + 255) / 256, 256, 0, s>>>(w, Nout, K, tcd.NP, tcd.KP, n_off, tcd.dst);
        YFV2_LAUNCH_CHECK();
        float* aff = tcd.dst + 2 * (size_t)tcd.NP * tcd.KP;
        // reuse the affine part of pack_pw_kernel: Nout*1 "weights" written to a scratch-free position is avoided by K=0
        pack_pw_kernel<<<(Nout + 255) / 256, 256, 0, s>>>(w, Nout, 0, g, b, m, v, bi, aff, tcd.NP, n_off, aff, aff + tcd.NP);
        YFV2_LAUNCH_CHECK();
    }
    return YFV2_OK;
}

int pack_dw(Cursor& cur, int Cn, int ksz, float* pack, cudaStream_t s) {
    const float* w = cur.params[cur.p];
    const float *g = cur.params[cur.p + 1], *b = cur.params[cur.p + 2], *m = cur.bn[2 * cur.b], *v = cur.bn[2 * cur.b + 1];
    cur.p += 3; cur.b += 1;
    if (!w || !g || !b || !m || !v) { set_error("pack_weights: null tensor pointer near param %d", cur.p); return YFV2_EINVAL; }
    const int KK = ksz * ksz, R = ksz == 3 ? 12 : 28;
    pack_dw_kernel<<<(Cn * KK + 255) / 256, 256, 0, s>>>(w, Cn, KK, R, g, b, m, v, pack);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

}  // namespace

extern "C" int yfv2_pack_weights(yfv2_plan* p, const float* const* params, const float* const* bn_running, void* packed,
                                 void* stream) {
    if (!p || !params || !bn_running || !packed) { set_error("pack_weights: null argument"); return YFV2_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    float* pk = (float*)packed;
    YFV2_CUDA(cudaMemsetAsync(pk, 0, p->pk_floats * sizeof(float), s));
    Cursor cur{params, bn_running};
    // backbone.first_conv (shufflenetv2.py:74-78)
    TRY(pack_pw(cur, true, false, 24, 27, pk + p->pk_stem, 24, 0, s, TcDst{pk + p->tk_stem, 32, 32}));
    for (int b = 0; b < kNumBlocks; ++b) {
        const int K = p->blk_K[b];
        float* d = pk + p->pk_block[b];
        const int pwn = pw_pack_floats(K, K), dwn = dw3_pack_floats(K);
        const int NPk = tc::tc_round(K, 16);
        const TcDst t1{pk + p->tk_blk[b][0], NPk, K}, t2{pk + p->tk_blk[b][1], NPk, K}, tp{pk + p->tk_blk[b][2], NPk, K};
        if (p->blk_stride[b] == 2) {
            // state_dict order: branch_main (pw1, dw, pw2) then branch_proj (dw, pw); pack order DWp|PWp|PW1|DW|PW2
            TRY(pack_pw(cur, true, false, K, K, d + dwn + pwn, K, 0, s, t1));
            TRY(pack_dw(cur, K, 3, d + dwn + 2 * pwn, s));
            TRY(pack_pw(cur, true, false, K, K, d + 2 * dwn + 2 * pwn, K, 0, s, t2));
            TRY(pack_dw(cur, K, 3, d, s));
            TRY(pack_pw(cur, true, false, K, K, d + dwn, K, 0, s, tp));
        } else {
            TRY(pack_pw(cur, true, false, K, K, d, K, 0, s, t1));
            TRY(pack_dw(cur, K, 3, d + pwn, s));
            TRY(pack_pw(cur, true, false, K, K, d + pwn + dwn, K, 0, s, t2));
        }
    }
    // fpn: conv1x1_2, conv1x1_3, cls_head_2, reg_head_2, reg_head_3, cls_head_3 (fpn.py:35-49 registration order)
    TRY(pack_pw(cur, true, false, kFpnDepth, 288, pk + p->pk_fpn2, kFpnDepth, 0, s, TcDst{pk + p->tk_fpn2, 80, 288}));
    TRY(pack_pw(cur, true, false, kFpnDepth, 192, pk + p->pk_fpn3, kFpnDepth, 0, s, TcDst{pk + p->tk_fpn3, 80, 192}));
    const int head_order[4] = {0, 1, 3, 2};     // pk_head index: 0 cls2, 1 reg2, 2 cls3, 3 reg3
    struct Pw2 { const float *w, *g, *b, *m, *v; } pw2[4];
    for (int i = 0; i < 4; ++i) {
        float* d = pk + p->pk_head[head_order[i]];
        const int dwn = dw5_pack_floats(kFpnDepth), pwn = pw_pack_floats(kFpnDepth, kFpnDepth);
        const int hi_ = head_order[i];
        TRY(pack_dw(cur, kFpnDepth, 5, d, s));
        TRY(pack_pw(cur, true, false, kFpnDepth, kFpnDepth, d + dwn, kFpnDepth, 0, s, TcDst{pk + p->tk_head[hi_][0], 80, 72}));
        TRY(pack_dw(cur, kFpnDepth, 5, d + dwn + pwn, s));
        pw2[hi_] = Pw2{cur.params[cur.p], cur.params[cur.p + 1], cur.params[cur.p + 2], cur.bn[2 * cur.b], cur.bn[2 * cur.b + 1]};
        TRY(pack_pw(cur, true, false, kFpnDepth, kFpnDepth, d + 2 * dwn + pwn, kFpnDepth, 0, s, TcDst{pk + p->tk_head[hi_][1], 80, 72}));
    }
    // output_reg_layers, output_obj_layers, output_cls_layers (detector.py:17-19); obj and cls share one pack
    // the chained output convs contract over the 80-column (72 real) feature tile of the head's second pointwise
    const bool tc_ok = true;
    const float* wo[3] = {cur.params[cur.p], cur.params[cur.p + 2], cur.params[cur.p + 4]};       // reg, obj, cls weights
    const float* bo[3] = {cur.params[cur.p + 1], cur.params[cur.p + 3], cur.params[cur.p + 5]};   // and biases
    TRY(pack_pw(cur, false, true, 4 * p->A, kFpnDepth, pk + p->pk_out_reg, round4(4 * p->A), 0, s,
                tc_ok ? TcDst{pk + p->tk_out_reg, 96, 80} : TcDst{nullptr, 0, 0}));
    const int Moc = round4(p->A + p->C);
    TRY(pack_pw(cur, false, true, p->A, kFpnDepth, pk + p->pk_out_oc, Moc, 0, s, tc_ok ? TcDst{pk + p->tk_out_oc, 96, 80} : TcDst{nullptr, 0, 0}));
    TRY(pack_pw(cur, false, true, p->C, kFpnDepth, pk + p->pk_out_oc, Moc, p->A, s, tc_ok ? TcDst{pk + p->tk_out_oc, 96, 80} : TcDst{nullptr, 0, 0}));
    // folded second halves of the four heads (cls heads feed obj|cls, reg heads feed reg)
    for (int h = 0; h < 4; ++h) {
        const Pw2& q = pw2[h];
        float* Fw = pk + p->pk_foldw[h];
        float* Fb = Fw + 96 * kFpnDepth;
        const int K = kFpnDepth;
        auto fold = [&](int which, int Mo, int n_off) -> int {
            if (!wo[which] || !bo[which]) { set_error("pack_weights: null output-layer tensor"); return YFV2_EINVAL; }
            fold_head_kernel<<<(Mo * K + 255) / 256, 256, 0, s>>>(wo[which], bo[which], Mo, q.w, q.g, q.b, q.m, q.v, K, n_off, Fw, Fb);
            YFV2_LAUNCH_CHECK();
            return YFV2_OK;
        };
        if (h & 1) { TRY(fold(0, 4 * p->A, 0)); }
        else { TRY(fold(1, p->A, 0)); TRY(fold(2, p->C, p->A)); }
        float* dst = pk + p->tk_fold[h];
        pack_tc_kernel<<<(96 * K + 255) / 256, 256, 0, s>>>(Fw, 96, K, 96, K, 0, dst);
        YFV2_LAUNCH_CHECK();
        fold_affine_kernel<<<1, 96, 0, s>>>(Fb, 96, dst + 2 * 96 * K);
        YFV2_LAUNCH_CHECK();
    }
    if (cur.p != YFV2_NUM_PARAMS || cur.b != YFV2_NUM_BN) {
        set_error("pack_weights: internal walk consumed %d params / %d BN layers", cur.p, cur.b);
        return YFV2_EINVAL;
    }
    return YFV2_OK;
}

namespace yfv2 {
bool stem2_supported(const StemArgs& a);
int launch_stem2(const StemArgs& a, cudaStream_t s);
int tc_launch_stem(const void* x, int is_u8, const Planes& out, const float* wpack, int N, int H, int W, cudaStream_t s);
int tc_launch_s1(int K, const Planes& P, const ChanTab& tin, const ChanTab& tout, const float* w1, const float* wdw, const float* w2,
                 int N, cudaStream_t s);
int tc_launch_s2(int K, const Planes& in, const Planes& out, const ChanTab& tin, const ChanTab& tout, const float* wdwp, const float* wp,
                 const float* w1, const float* wdwm, const float* w2, int N, cudaStream_t s);
int tc_launch_pw(int kind, const Planes& A, const ChanTab& ta, const Planes& B, const ChanTab& tb, const Planes& out, const ChanTab& tout,
                 const float* wpack, int N, cudaStream_t s);
bool tc_dws2c_supported(int K, const Planes& in, const Planes& out, int N, int* imgs_out, int* G_out, size_t* bytes_out);
int tc_launch_dws2c(int K, int nbranch, const Planes* in, const ChanTab* tin, const Planes* out, const ChanTab* tout,
                    const float* const* wdw, const float* const* wpw, int N, cudaStream_t s);
int tc_launch_dwpw96(int stride, int nbranch, const Planes* in, const ChanTab* tin, const Planes* out, const ChanTab* tout,
                     const float* const* wdw, const float* const* wpw, int N, cudaStream_t s);
int tc_launch_heads(int half, const Planes& sIn, const Planes& tcls, const Planes& treg, const float* const wdw[2], const float* const wpw[2],
                    float* reg, float* obj, float* cls, int A, int C, int N, cudaStream_t s);
// k_blk.cu
bool blk_s1_chainable(int K, int H, int W);
int blk_launch_s1(int K, const Planes& P, int nblk, const ChanTab* tin, const ChanTab* tout, const float* const* w1,
                  const float* const* wdw, const float* const* w2, int N, cudaStream_t s, int* done);
int blk_launch_s2(int K, const Planes& in, const Planes& out, const ChanTab& tin, const ChanTab& tout, const float* wdwp, const float* wp,
                  const float* w1, const float* wdwm, const float* w2, int N, cudaStream_t s);
int blk_launch_pw(int kind, const Planes& A, const ChanTab& ta, const Planes& B, const ChanTab& tb, const Planes& out, const ChanTab& tout,
                  const float* wpack, int N, cudaStream_t s);
// k_tail.cu
int tail_launch_s1(const Planes& P, int nblk, const ChanTab* tin, const ChanTab* tout, const float* const* w1,
                   const float* const* wdw, const float* const* w2, int N, cudaStream_t s, int* done);
}

namespace {
// Runs fused stages [first, last) of the forward; each stage is exactly one kernel launch.
int forward_impl(yfv2_plan* p, const void* x, int is_u8, const void* packed, float* const preds[6], void* workspace,
                 int first, int last, cudaStream_t s) {
    if (!p || !x || !packed || !preds || !workspace) { set_error("forward: null argument"); return YFV2_EINVAL; }
    for (int i = 0; i < 6; ++i) if (!preds[i]) { set_error("forward: null output %d", i); return YFV2_EINVAL; }
    if (last < 0) last = p->n_stages;
    if (first < 0 || last > p->n_stages || first > last) { set_error("forward: bad stage range [%d,%d)", first, last); return YFV2_EINVAL; }
    float* ws = (float*)workspace;
    const float* pk = (const float*)packed;
    pdl_reset();                                            // the first kernel of this call is serialized normally
    if (p->ws_zeroed != workspace) {
        // the zero frames around every plane are never written afterwards (kernels store interior pixels only)
        YFV2_CUDA(cudaMemsetAsync(workspace, 0, p->ws_floats * sizeof(float), s));
        p->ws_zeroed = workspace;
    }
    ChanTab ident;
    for (int i = 0; i < kMaxCh; ++i) ident.c[i] = (unsigned short)i;
    struct { Planes c3, c2, s3, s2; ChanTab t3, t2; } f;
    f.c3 = pool_planes(p, ws, 3); f.t3 = p->c3;
    f.c2 = pool_planes(p, ws, 2); f.t2 = p->c2;
    f.s3 = flat_planes(p, ws, p->off_s3, 3);
    f.s2 = flat_planes(p, ws, p->off_s2, 2);
    for (int si = first; si < last; ++si) {
        const yfv2_plan::Stage& st = p->stages[si];
        const int b = st.a;
        switch (st.kind) {
        case 0: {
            static const bool stem_tc = getenv("YFV2_STEM_TC") != nullptr;
            static const bool stem_ffma = getenv("YFV2_STEM_FFMA") != nullptr;   // round-1 FFMA2 stem, kept for A/B runs
            if (!stem_tc) {
                StemArgs a{x, is_u8, p->N, p->H, p->W, pool_planes(p, ws, 0), pk + p->pk_stem};
                // default: the tensor-core strip walk (k_stem2.cu); the register-tiled FFMA2 direct convolution (k_stem.cu)
                // takes over for inputs whose base address is not 16-byte (uint8: 4-byte) aligned
                if (!stem_ffma && stem2_supported(a)) { TRY(launch_stem2(a, s)); }
                else { TRY(launch_stem(a, s)); }
            } else {
                TRY(tc_launch_stem(x, is_u8, pool_planes(p, ws, 0), pk + p->tk_stem, p->N, p->H, p->W, s));
            }
        } break;
        case 15: {
            const int lv = st.a, half = st.b;
            const Planes sIn = lv ? f.s3 : f.s2;
            const Planes t_cls = flat_planes(p, ws, p->off_t[2 * lv], 2 + lv);
            const Planes t_reg = flat_planes(p, ws, p->off_t[2 * lv + 1], 2 + lv);
            const float* w_cls = pk + p->pk_head[2 * lv];
            const float* w_reg = pk + p->pk_head[2 * lv + 1];
            // DW pack = first 72*28 floats of each half of the head pack
            const size_t half_off = (size_t)half * (dw5_pack_floats(kFpnDepth) + pw_pack_floats(kFpnDepth, kFpnDepth));
            const float* wdw[2] = {w_cls + half_off, w_reg + half_off};
            const float* wpw[2] = {half ? pk + p->tk_fold[2 * lv] : pk + p->tk_head[2 * lv][0],
                                   half ? pk + p->tk_fold[2 * lv + 1] : pk + p->tk_head[2 * lv + 1][0]};
            TRY(tc_launch_heads(half, sIn, t_cls, t_reg, wdw, wpw, preds[3 * lv], preds[3 * lv + 1],
                                preds[3 * lv + 2], p->A, p->C, p->N, s));
        } break;
        case 10: {   // fused stride-1 blocks; consecutive blocks of a stage inside [first, last) share one launch when
                     // an image's intermediate fits in shared memory (k_blk.cu)
            const int K = p->blk_K[b];
            static const bool old_s1 = getenv("YFV2_S1_OLD") != nullptr;          // round-1 kernel, kept for A/B runs
            if (old_s1) {
                const float* d = pk + p->pk_block[b];
                TRY(tc_launch_s1(K, pool_planes(p, ws, p->blk_res[b]), p->tin[b], p->tout[b], pk + p->tk_blk[b][0],
                                 d + pw_pack_floats(K, K), pk + p->tk_blk[b][1], p->N, s));
                break;
            }
            int nb = 1;
            while (si + nb < last && p->stages[si + nb].group == st.group) ++nb;
            const float *w1[8], *w2[8], *wd[8];
            if (nb > 7) nb = 7;
            for (int j = 0; j < nb; ++j) {
                const int bj = b + j;
                w1[j] = pk + p->tk_blk[bj][0]; w2[j] = pk + p->tk_blk[bj][1];
                wd[j] = pk + p->pk_block[bj] + pw_pack_floats(K, K);
            }
            int done = 1;
            TRY(blk_launch_s1(K, pool_planes(p, ws, p->blk_res[b]), nb, &p->tin[b], &p->tout[b], w1, wd, w2, p->N, s, &done));
            si += done - 1;
        } break;
        case 11: {   // tc fused stride-2 block
            const int K = p->blk_K[b];
            const float* d = pk + p->pk_block[b];
            const int dwn = dw3_pack_floats(K), pwn = pw_pack_floats(K, K);
            static const bool old_s2 = getenv("YFV2_S2_OLD") != nullptr;          // round-1 kernel, kept for A/B runs
            auto s2 = old_s2 ? tc_launch_s2 : blk_launch_s2;
            TRY(s2(K, pool_planes(p, ws, p->blk_res[b] - 1), pool_planes(p, ws, p->blk_res[b]), p->tin[b], p->tout[b],
                   d, pk + p->tk_blk[b][2], pk + p->tk_blk[b][0], d + dwn + 2 * pwn, pk + p->tk_blk[b][1], p->N, s));
        } break;
        case 16: {   // K=96 stride-1 blocks on a map of at most 128 pixels: one chained launch (k_tail.cu)
            int nb = 1;
            while (si + nb < last && p->stages[si + nb].group == st.group) ++nb;
            const float *w1[4], *w2[4], *wd[4];
            if (nb > 4) nb = 4;
            for (int j = 0; j < nb; ++j) {
                const int bj = b + j;
                w1[j] = pk + p->tk_blk[bj][0]; w2[j] = pk + p->tk_blk[bj][1];
                wd[j] = pk + p->pk_block[bj] + pw_pack_floats(96, 96);
            }
            int done = 1;
            TRY(tail_launch_s1(pool_planes(p, ws, p->blk_res[b]), nb, &p->tin[b], &p->tout[b], w1, wd, w2, p->N, s, &done));
            si += done - 1;
        } break;
        case 12: case 13: {   // K=96 blocks (and the K=48 stride-2 block): pw1 (global -> scratch planes), then dw3x3 -> pw2 (+ proj branch when stride 2)
            const int stride = p->blk_stride[b], K = p->blk_K[b];
            const Planes out = pool_planes(p, ws, p->blk_res[b]);
            const Planes in = stride == 2 ? pool_planes(p, ws, p->blk_res[b] - 1) : out;
            ChanTab t96;
            // scratch planes live behind the regular planes of the INPUT resolution's pool
            const int base = K == 48 ? 72 : (stride == 2 ? 144 : 288);
            for (int i = 0; i < K; ++i) t96.c[i] = (unsigned short)(base + i);
            static const bool old_pw = getenv("YFV2_PW_OLD") != nullptr;          // round-1 pointwise kernel, kept for A/B runs
            if (st.kind == 12) {
                if (K == 48) { TRY(blk_launch_pw(3, in, p->tin[b], in, p->tin[b], in, t96, pk + p->tk_blk[b][0], p->N, s)); }
                else if (old_pw) { TRY(tc_launch_pw(0, in, p->tin[b], in, p->tin[b], in, t96, pk + p->tk_blk[b][0], p->N, s)); }
                else { TRY(blk_launch_pw(0, in, p->tin[b], in, p->tin[b], in, t96, pk + p->tk_blk[b][0], p->N, s)); }
            } else {
                const float* d = pk + p->pk_block[b];
                const int dwn = dw3_pack_floats(K), pwn = pw_pack_floats(K, K);
                if (stride == 1) {
                    const Planes ins[1] = {in}; const ChanTab tins[1] = {t96}; const Planes outs[1] = {out}; const ChanTab touts[1] = {p->tout[b]};
                    const float* wdw[1] = {d + pwn}; const float* wpw[1] = {pk + p->tk_blk[b][1]};
                    TRY(tc_launch_dwpw96(1, 1, ins, tins, outs, touts, wdw, wpw, p->N, s));
                } else {
                    ChanTab tmain;
                    for (int i = 0; i < K; ++i) tmain.c[i] = p->tout[b].c[K + i];
                    const Planes ins[2] = {in, in}; const ChanTab tins[2] = {p->tin[b], t96};
                    const Planes outs[2] = {out, out}; const ChanTab touts[2] = {p->tout[b], tmain};
                    const float* wdw[2] = {d, d + dwn + 2 * pwn}; const float* wpw[2] = {pk + p->tk_blk[b][2], pk + p->tk_blk[b][1]};
                    static const bool old_s2 = getenv("YFV2_S2_96_OLD") != nullptr;       // round-1 banded kernel, kept for A/B runs
                    int im_ = 0, g_ = 0; size_t by_ = 0;
                    if ((K == 48 || !old_s2) && tc_dws2c_supported(K, in, out, p->N, &im_, &g_, &by_)) { TRY(tc_launch_dws2c(K, 2, ins, tins, outs, touts, wdw, wpw, p->N, s)); }
                    else if (K == 96) { TRY(tc_launch_dwpw96(2, 2, ins, tins, outs, touts, wdw, wpw, p->N, s)); }
                    else { set_error("forward: the K=48 stride-2 split needs an output map of at most 512 pixels"); return YFV2_EUNSUPPORTED; }
                }
            }
        } break;
        case 14: {   // tc FPN reducers
            static const bool old_pw = getenv("YFV2_PW_OLD") != nullptr;
            auto pw = old_pw ? tc_launch_pw : blk_launch_pw;
            if (st.a == 1) { TRY(pw(1, f.c3, f.t3, f.c3, f.t3, f.s3, ident, pk + p->tk_fpn3, p->N, s)); }
            else { TRY(pw(2, f.c3, f.t3, f.c2, f.t2, f.s2, ident, pk + p->tk_fpn2, p->N, s)); }
        } break;
        default: set_error("forward: unknown stage kind %d", st.kind); return YFV2_EINVAL;
        }
    }
    return YFV2_OK;
}
}  // namespace

extern "C" int yfv2_forward_range(yfv2_plan* p, const void* x, int is_u8, const void* packed, float* const preds[6],
                                  void* workspace, int first, int last, void* stream) {
    return forward_impl(p, x, is_u8, packed, preds, workspace, first, last, (cudaStream_t)stream);
}

extern "C" int yfv2_forward(yfv2_plan* p, const float* x, const void* packed, float* const preds[6], void* workspace,
                            void* stream) {
    return forward_impl(p, x, 0, packed, preds, workspace, 0, -1, (cudaStream_t)stream);
}

extern "C" int yfv2_forward_u8(yfv2_plan* p, const uint8_t* x, const void* packed, float* const preds[6], void* workspace,
                               void* stream) {
    return forward_impl(p, x, 1, packed, preds, workspace, 0, -1, (cudaStream_t)stream);
}

// ---- whole step with host buffers ---------------------------------------------------------------------------
namespace {
struct DetectLayout {
    size_t off_x, off_pred[6], off_out, off_counts, total;
};
DetectLayout detect_layout(const yfv2_plan* p, int max_det) {
    DetectLayout L;
    size_t off = align_up(p->ws_floats * sizeof(float), 256);
    L.off_x = off; off += align_up((size_t)p->N * 3 * p->H * p->W, 256);
    for (int lv = 0; lv < 2; ++lv) {
        const size_t hw = (size_t)p->h[2 + lv] * p->w[2 + lv];
        const int ch[3] = {4 * p->A, p->A, p->C};
        for (int k = 0; k < 3; ++k) { L.off_pred[3 * lv + k] = off; off += align_up((size_t)p->N * ch[k] * hw * sizeof(float), 256); }
    }
    L.off_out = off; off += align_up((size_t)p->N * max_det * 6 * sizeof(float), 256);
    L.off_counts = off; off += align_up((size_t)p->N * sizeof(int), 256);
    L.total = off;
    return L;
}
}  // namespace

extern "C" size_t yfv2_detect_workspace_bytes(const yfv2_plan* p, int max_det) {
    if (!p || max_det <= 0) return 0;
    return detect_layout(p, max_det).total;
}

extern "C" int yfv2_detect_u8_host(yfv2_plan* p, const uint8_t* x_host, const void* packed, const double* anchors_host,
                                   float conf_thres, double iou_thres, int max_det, float* out_host, int* counts_host,
                                   void* workspace, void* stream) {
    if (!p || !x_host || !packed || !anchors_host || !out_host || !counts_host || !workspace || max_det <= 0) {
        set_error("detect_u8_host: null argument");
        return YFV2_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const DetectLayout L = detect_layout(p, max_det);
    unsigned char* ws = (unsigned char*)workspace;
    uint8_t* x_dev = ws + L.off_x;
    float* preds[6];
    for (int i = 0; i < 6; ++i) preds[i] = (float*)(ws + L.off_pred[i]);
    float* out_dev = (float*)(ws + L.off_out);
    int* counts_dev = (int*)(ws + L.off_counts);
    YFV2_CUDA(cudaMemcpyAsync(x_dev, x_host, (size_t)p->N * 3 * p->H * p->W, cudaMemcpyHostToDevice, s));
    struct PdlOff { PdlOff() { g_pdl_call = false; } ~PdlOff() { g_pdl_call = true; } } pdl_off_guard;
    TRY(forward_impl(p, x_dev, 1, packed, preds, ws, 0, -1, s));
    TRY(yfv2_decode_nms(preds, p->N, p->H, p->W, p->A, p->C, anchors_host, conf_thres, iou_thres, nullptr, 0, max_det,
                        4096.0f, out_dev, counts_dev, nullptr, nullptr, stream));
    YFV2_CUDA(cudaMemcpyAsync(out_host, out_dev, (size_t)p->N * max_det * 6 * sizeof(float), cudaMemcpyDeviceToHost, s));
    YFV2_CUDA(cudaMemcpyAsync(counts_host, counts_dev, (size_t)p->N * sizeof(int), cudaMemcpyDeviceToHost, s));
    return YFV2_OK;
}

// ---- debug: dense NCHW copy of an intermediate tensor ------------------------------------------------------
extern "C" int yfv2_debug_gather(const yfv2_plan* p, const void* workspace, int which, float* out, int* dims4, void* stream) {
    if (!p || !workspace || !dims4) { set_error("debug_gather: null argument"); return YFV2_EINVAL; }
    float* ws = (float*)workspace;
    Planes P; ChanTab tab; int Cn;
    for (int i = 0; i < kMaxCh; ++i) tab.c[i] = (unsigned short)i;
    if (which == 0) { P = pool_planes(p, ws, 0); Cn = 24; }
    else if (which >= 1 && which <= kNumBlocks) { const int b = which - 1; P = pool_planes(p, ws, p->blk_res[b]); Cn = 2 * p->blk_K[b]; tab = p->logical[b]; }
    else if (which == 17) { P = flat_planes(p, ws, p->off_s2, 2); Cn = kFpnDepth; }
    else if (which == 18) { P = flat_planes(p, ws, p->off_s3, 3); Cn = kFpnDepth; }
    else if (which >= 19 && which <= 22) { const int i = which - 19; P = flat_planes(p, ws, p->off_t[i], i < 2 ? 2 : 3); Cn = kFpnDepth; }
    else { set_error("debug_gather: unknown tensor id %d", which); return YFV2_EINVAL; }
    dims4[0] = p->N; dims4[1] = Cn; dims4[2] = P.H; dims4[3] = P.W;
    if (!out) return YFV2_OK;
    const long long total = (long long)p->N * Cn * P.H * P.W;
    gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, tab, Cn, out, total);
    YFV2_LAUNCH_CHECK();
    return YFV2_OK;
}
