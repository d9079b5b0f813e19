// tcgen05 / TMEM / TMA implicit-GEMM convolution (stride 1, 1x1 or 3x3, bf16 NHWC, fp32 accumulate).
//
// GEMM view: D[M = 128 output pixels][N = BN output channels] += A[M][K] * B[N][K]^T with
// K = taps * Cin.  One k-block = (filter tap, 64-channel chunk):
//   A tile : a TMA box {64 ch, TW, TH, NB} of the NHWC activation tensor whose origin is shifted by the
//            tap offset (r - pad, s - pad); out-of-bounds elements (the conv zero padding and ragged
//            tile edges) are zero-filled by the TMA unit.  In shared memory the box is 128 rows
//            (pixels, w fastest) x 128 bytes, SWIZZLE_128B == the canonical K-major UMMA operand.
//   B tile : a TMA box {64 ch, BN, 1} of the packed weights [tap][Cout][Cin] (K-major as well).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warps 2..5 = epilogue (TMEM -> registers -> bias/activation -> bf16 -> global); warp 2 owns the
// TMEM allocation.  A second activation tensor (x1) realises torch.cat([x0, x1], 1) without
// materialising it: channel chunks beyond x0's come from x1's tensor map.
#include "tc_common.cuh"
#include <cudaTypedefs.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace ssg {
namespace tc {

constexpr int BM = 128;          // pixels per tile == UMMA M
constexpr int BK = 64;           // bf16 channels per k-block (128 bytes == swizzle span)
constexpr int A_BYTES = BM * BK * 2;
constexpr int NUM_THREADS = 192;

// One output-parity class of a (possibly strided / transposed) convolution: the taps that contribute to it, as
// source-pixel offsets (added to in_mul * tile pixel) and the index of the weight tap they multiply.
struct TapClass {
    int ntaps;
    int8_t dy[9], dx[9], wt[9];
    int8_t add_y, add_x;         // destination pixel = out_mul * tile pixel + add
    int8_t pad_[3];
};

struct ConvTcParams {
    bf16* y;
    const float* bias;
    int N, H, W;                 // tile-space dims: the pixel grid the GEMM rows enumerate
    int out_H, out_W;            // spatial dims of y
    int out_mul;                 // 1; 2 for the data gradient of a stride-2 convolution (one class per parity)
    int in_mul;                  // 1; 2 for a stride-2 forward (the tensor map traverses with the same stride)
    int cout;                    // output channels as stored (row stride of y)
    int bias_n;                  // number of valid bias entries (real channels; padding channels get none)
    int tw_log2, th_log2;        // tile = TW x TH x NB pixels (product 128)
    int tiles_x, tiles_y;
    int chunks0, chunks1;        // 64-channel chunks taken from x0 / x1
    int act;
    float slope;
    const bf16* mask;            // optional (data gradient): a tensor shaped like y holding the post-activation OUTPUT of the layer that
    int mask_act;                // produced this convolution's input; the epilogue multiplies by act'(mask), i.e. the producer's
    float mask_slope;            // activation backward (models_seg_gan.py:52-57 LeakyReLU) rides along instead of a separate pass
    int ncls;                    // number of classes
    int merge;                   // 1: ONE CTA computes all classes of its tile (one TMEM accumulator each) instead of one CTA per
                                 // class: the stride-2 data gradient's four parity classes share their dy boxes (L2) and the CTA's
                                 // fixed costs, and the nine k-blocks fill one pipeline (ncu: 8 % tensor activity, dy read 4 x before)
    TapClass cls[4];             // blockIdx.z selects the class (merge == 0)
};

// ST: ring depth.  0 = the default, 3 stages (round 1 used 4 below BN = 128: with 3 a third CTA fits per SM, and these CTAs are
// short -- 9 k-blocks for a stride-2 layer with Cin = 64 -- so residency beats depth: D block 1 forward 0.182 -> 0.163 ms, data
// gradient 0.311 -> 0.284 ms); SSG_PLAIN_STAGES=2 / 3 overrides (A/B switch)
template <int BN, int ST = 0>
struct SmemLayout {
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = ST > 0 ? ST : 3;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // + barriers + alignment slack
    static constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN, int ST>
__global__ void __launch_bounds__(NUM_THREADS) conv_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                                   const __grid_constant__ CUtensorMap tmA1,
                                                                   const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
    using L = SmemLayout<BN, ST>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + L::STAGES;
    uint64_t* tmem_full_bar = empty_bar + L::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TW = 1 << p.tw_log2, TH = 1 << p.th_log2;
    // tile -> (image group, tile row, tile col)
    int tile = blockIdx.x;
    const int tx = tile % p.tiles_x; tile /= p.tiles_x;
    const int ty = tile % p.tiles_y;
    const int tn = tile / p.tiles_y;
    const int w0 = tx * TW, h0 = ty * TH, img0 = tn * (BM >> (p.tw_log2 + p.th_log2));
    const int n0 = blockIdx.y * BN;
    const int chunks = p.chunks0 + p.chunks1;
    const int cls0 = p.merge ? 0 : (int)blockIdx.z, ncls_here = p.merge ? p.ncls : 1;
    const uint32_t tmem_cols = p.merge ? 4u * L::TMEM_COLS : L::TMEM_COLS;

    if (threadIdx.x == 0) {
        for (int i = 0; i < L::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tmA0);
            tma_prefetch_desc(&tmB);
            int g = 0;                               // k-block counter over all classes of this CTA (ring position)
            for (int ci = 0; ci < ncls_here; ++ci) {
                const TapClass& tcl = p.cls[cls0 + ci];
                const int num_kb = tcl.ntaps * chunks;
                for (int kb = 0; kb < num_kb; ++kb, ++g) {
                    const int stage = g % L::STAGES;
                    const uint32_t phase = (g / L::STAGES) & 1;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                    const int tap = kb / chunks, ch = kb - tap * chunks;
                    const int cx = p.in_mul * w0 + tcl.dx[tap], cy = p.in_mul * h0 + tcl.dy[tap];
                    uint8_t* sa = smem + stage * L::STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    if (ch < p.chunks0) tma_load_4d(sa, &tmA0, ch * BK, cx, cy, img0, &full_bar[stage]);
                    else tma_load_4d(sa, &tmA1, (ch - p.chunks0) * BK, cx, cy, img0, &full_bar[stage]);
                    tma_load_3d(sb, &tmB, ch * BK, n0, tcl.wt[tap], &full_bar[stage]);
                }
            }
        }
    } else if (warp == 1) {
        // whole warp walks the (uniform) loop; the elected lane issues tcgen05.mma / commit (uniform-datapath issue sequence)
        constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
        const bool leader = elect_one();
        const uint32_t hi = desc_hi(1024, 2);
        int g = 0;
        for (int ci = 0; ci < ncls_here; ++ci) {
            const int num_kb = p.cls[cls0 + ci].ntaps * chunks;
            const uint32_t d_tmem = tmem_base + (uint32_t)(ci * (int)L::TMEM_COLS);      // one accumulator per class
            for (int kb = 0; kb < num_kb; ++kb, ++g) {
                const int stage = g % L::STAGES;
                const uint32_t phase = (g / L::STAGES) & 1;
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
                const uint32_t a_lo = desc_lo(sa, 16), b_lo = desc_lo(sa + A_BYTES, 16);
                if (leader) {
                    umma_bf16_lohi(d_tmem, a_lo, hi, b_lo, hi, idesc, (uint32_t)kb);
#pragma unroll
                    for (int k = 1; k < BK / 16; ++k)   // +32 bytes (>>4 == 2) per 16-element K step inside the swizzle atom
                        umma_bf16_lohi(d_tmem, a_lo + (uint32_t)(2 * k), hi, b_lo + (uint32_t)(2 * k), hi, idesc, 1u);
                    umma_commit(&empty_bar[stage]);     // frees the smem slot once these MMAs have read it
                }
                __syncwarp();
            }
        }
        if (leader) umma_commit(tmem_full_bar);     // accumulators complete
        __syncwarp();
    } else {
        // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) ----
        const int q = warp & 3;
        const int m = q * 32 + lane;                // accumulator row == pixel index inside the tile
        const int twi = m & (TW - 1), thi = (m >> p.tw_log2) & (TH - 1), nbi = m >> (p.tw_log2 + p.th_log2);
        const int tx_ = w0 + twi, ty_ = h0 + thi, on = img0 + nbi;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const bool vec_ok = (p.cout % 8 == 0);
        const bool has_bias = p.bias != nullptr;
        // act(v) = max(v, v * neg): neg = 1 (identity), 0 (ReLU) or the LeakyReLU slope -- branch-free
        const float neg = p.act == SSG_ACT_RELU ? 0.f : (p.act == SSG_ACT_LEAKY ? p.slope : 1.f);
#pragma unroll 1
        for (int ci = 0; ci < ncls_here; ++ci) {
        const TapClass& tcl = p.cls[cls0 + ci];
        const int ox = p.out_mul * tx_ + tcl.add_x, oy = p.out_mul * ty_ + tcl.add_y;
        const bool row_ok = tx_ < p.W && ty_ < p.H && on < p.N && ox < p.out_W && oy < p.out_H;
        const long long row_off = (((long long)on * p.out_H + oy) * p.out_W + ox) * p.cout + n0;
        bf16* yrow = p.y + row_off;
        const bf16* mrow = p.mask ? p.mask + row_off : nullptr;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            if (n0 + c0 >= p.cout) break;
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ci * (int)L::TMEM_COLS + c0), v);
            tmem_ld_wait();
            if (!row_ok) continue;
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
            if (has_bias) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (n0 + c0 + j < p.bias_n) f[j] += __ldg(p.bias + n0 + c0 + j);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], f[j] * neg);
            if (mrow != nullptr) {
                if (vec_ok && n0 + c0 + 16 <= p.cout) {          // two 16-byte loads of the producer's output row
                    Vec<bf16> m0, m1;
                    m0.load(mrow + c0);
                    m1.load(mrow + c0 + 8);
                    float mv[16];
                    m0.get(mv);
                    m1.get(mv + 8);
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] *= act_grad_from_out(mv[j], p.mask_act, p.mask_slope);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + c0 + j < p.cout) f[j] *= act_grad_from_out(__bfloat162float(mrow[c0 + j]), p.mask_act, p.mask_slope);
                }
            }
            if (vec_ok && n0 + c0 + 16 <= p.cout) {
                Vec<bf16> o;
                o.set(f); o.store(yrow + c0);
                o.set(f + 8); o.store(yrow + c0 + 8);
            } else if (vec_ok && n0 + c0 + 8 <= p.cout) {
                Vec<bf16> o;
                o.set(f); o.store(yrow + c0);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (n0 + c0 + j < p.cout) yrow[c0 + j] = __float2bfloat16_rn(f[j]);
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// ---- host side ------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

// bf16 tensor map; dims/box innermost first; strides in BYTES for dims 1..rank-1
int run_conv_halo(const void* x0, int c0, const void* x1, int c1, const void* w_packed, int w_taps, const float* bias, int bias_n,
                  void* y, int n, int h, int w, int gemm_n, int ksize, int flip, int act, float slope, double* stats, cudaStream_t st,
                  int accumulate = 0, void* y1 = nullptr, int split_c = 0);
int run_wgrad_halo(const void* x0, int c0, const void* x1, int c1, const void* dy, int cout_s, float* dw, int cout_real, int cin_real,
                   int n, int h, int w, cudaStream_t st);
int run_wgrad_thin(const void* x, const void* dy, int cout_s, float* dw, int cout_real, int cin_real, int n, int h, int w,
                   cudaStream_t st);
bool dgrad_s2_halo_supported(int h, int w, int cin, int cout);
int run_dgrad_s2_halo(const void* dy, int cout, const void* w_packed, void* dx, int n, int h, int w, int cin, const void* mask,
                      int mask_act, float mask_slope, cudaStream_t st);
int encode_bf16_map(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, CUtensorMapSwizzle swizzle, const uint32_t* elem_strides);

static bool use_halo_kernel() {
    static const bool v1 = getenv("SSG_CONV_V1") != nullptr;    // debugging / A-B switch: force the per-tap kernel
    return !v1;
}

int encode_bf16_map(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, CUtensorMapSwizzle swizzle, const uint32_t* elem_strides) {
    auto enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return SSG_ERR_CUDA; }
    uint32_t estr[5] = {1, 1, 1, 1, 1};
    if (elem_strides) for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return SSG_ERR_CUDA; }
    return SSG_OK;
}

static int ilog2_ceil(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

template <int BN, int ST>
static int launch_fwd_st(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvTcParams& p, int m_tiles,
                         int n_classes, cudaStream_t st) {
    using L = SmemLayout<BN, ST>;
    static bool attr_set = false;
    if (!attr_set) {
        SSG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_fwd_kernel<BN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    dim3 grid((unsigned)m_tiles, (unsigned)((p.cout + BN - 1) / BN), (unsigned)(p.merge ? 1 : n_classes));
    conv_tc_fwd_kernel<BN, ST><<<grid, NUM_THREADS, L::TOTAL, st>>>(a0, a1, b, p);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

template <int BN>
static int launch_fwd(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvTcParams& p, int m_tiles,
                      int n_classes, cudaStream_t st) {
    static const int st_env = getenv("SSG_PLAIN_STAGES") ? atoi(getenv("SSG_PLAIN_STAGES")) : 0;
    if (st_env == 2) return launch_fwd_st<BN, 2>(a0, a1, b, p, m_tiles, n_classes, st);
    if (st_env == 3) return launch_fwd_st<BN, 3>(a0, a1, b, p, m_tiles, n_classes, st);
    return launch_fwd_st<BN, 0>(a0, a1, b, p, m_tiles, n_classes, st);
}

}  // namespace tc
}  // namespace ssg
using namespace ssg;
using namespace ssg::tc;

namespace ssg {
namespace tc {

static void pick_tile(int h, int w, int& twl, int& thl) {
    twl = ilog2_ceil(w); if (twl > 7) twl = 7;
    thl = ilog2_ceil(h); if (thl > 7 - twl) thl = 7 - twl;
}

// NHWC activation map: box = {64 ch, TW, TH, NB} output pixels; with in_mul == 2 the box spans 2TW x 2TH source
// pixels traversed with element stride 2 (the stride-2 convolution's gather).
static int encode_act_map(CUtensorMap* m, const void* ptr, int c, int n, int h, int w, int twl, int thl, int in_mul) {
    const int TW = 1 << twl, TH = 1 << thl, NB = BM / (TW * TH);
    uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    uint64_t str[3] = {(uint64_t)c * 2, (uint64_t)w * c * 2, (uint64_t)h * w * c * 2};
    uint32_t box[4] = {64, (uint32_t)(TW * in_mul), (uint32_t)(TH * in_mul), (uint32_t)NB};
    uint32_t es[4] = {1, (uint32_t)in_mul, (uint32_t)in_mul, 1};
    return encode_bf16_map(m, ptr, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, es);
}

// Shared driver of the forward / data-gradient launches.  (sh, sw): source spatial dims; (th_, tw_): tile-space dims;
// (oh, ow): destination dims; gemm_n: destination channels; weights are [taps][gemm_n][c0 + c1] bf16.
static int run_conv(const void* x0, int c0, const void* x1, int c1, const void* w_packed, int w_taps, const float* bias, int bias_n,
                    void* y, int n, int sh, int sw, int th_, int tw_, int oh, int ow, int gemm_n, int in_mul, int out_mul,
                    const TapClass* cls, int n_classes, int act, float slope, cudaStream_t st, const void* mask = nullptr,
                    int mask_act = 0, float mask_slope = 0.f) {
    int twl, thl;
    pick_tile(th_, tw_, twl, thl);
    const int TW = 1 << twl, TH = 1 << thl, NB = BM / (TW * TH);
    ConvTcParams p;
    memset(&p, 0, sizeof(p));
    p.y = (bf16*)y; p.bias = bias; p.bias_n = bias_n; p.N = n; p.H = th_; p.W = tw_; p.out_H = oh; p.out_W = ow; p.out_mul = out_mul; p.in_mul = in_mul;
    p.cout = gemm_n; p.tw_log2 = twl; p.th_log2 = thl;
    p.tiles_x = (tw_ + TW - 1) / TW; p.tiles_y = (th_ + TH - 1) / TH;
    p.chunks0 = (c0 + 63) / 64; p.chunks1 = (c1 + 63) / 64; p.act = act; p.slope = slope;
    p.mask = (const bf16*)mask; p.mask_act = mask_act; p.mask_slope = mask_slope;
    for (int i = 0; i < n_classes; ++i) p.cls[i] = cls[i];
    static const bool no_merge = getenv("SSG_S2_MERGE") != nullptr && atoi(getenv("SSG_S2_MERGE")) == 0;      // A/B switch
    p.ncls = n_classes;
    p.merge = (n_classes > 1 && !no_merge) ? 1 : 0;
    const int tiles_n = (n + NB - 1) / NB;
    CUtensorMap ma0, ma1, mb;
    int rc = encode_act_map(&ma0, x0, c0, n, sh, sw, twl, thl, in_mul);
    if (rc) return rc;
    ma1 = ma0;
    if (c1 > 0) {
        rc = encode_act_map(&ma1, x1, c1, n, sh, sw, twl, thl, in_mul);
        if (rc) return rc;
    }
    const int cin = c0 + c1;
    const int BN = gemm_n >= 128 ? 128 : (gemm_n > 16 ? 64 : 16);
    {
        uint64_t dims[3] = {(uint64_t)cin, (uint64_t)gemm_n, (uint64_t)w_taps};
        uint64_t str[2] = {(uint64_t)cin * 2, (uint64_t)gemm_n * cin * 2};
        uint32_t box[3] = {64, (uint32_t)BN, 1};
        rc = encode_bf16_map(&mb, w_packed, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, nullptr);
        if (rc) return rc;
    }
    const int m_tiles = p.tiles_x * p.tiles_y * tiles_n;
    if (BN == 128) return launch_fwd<128>(ma0, ma1, mb, p, m_tiles, n_classes, st);
    if (BN == 64) return launch_fwd<64>(ma0, ma1, mb, p, m_tiles, n_classes, st);
    return launch_fwd<16>(ma0, ma1, mb, p, m_tiles, n_classes, st);
}

}  // namespace tc
}  // namespace ssg

extern "C" {

int ssg_conv2d_fwd_tc(const void* x0, int c0, const void* x1, int c1, const void* w_packed, const float* bias, int bias_n, void* y,
                      int n, int h, int w, int cout, int ksize, int stride, int pad, int act, float slope, double* stats,
                      ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && cout > 0 && cout % 8 == 0 && c0 > 0 && c0 % 8 == 0 && c1 >= 0 && c1 % 8 == 0 &&
                      (c1 == 0 || c0 % 64 == 0),
                  "conv2d_fwd_tc: stored channel counts must be multiples of 8, c0 of 64 when x1 is given (c0=%d c1=%d cout=%d)", c0, c1,
                  cout);
    SSG_CHECK_ARG((ksize == 1 || ksize == 3) && pad >= 0 && pad < ksize && (stride == 1 || stride == 2),
                  "conv2d_fwd_tc: kernel 1 or 3, stride 1 or 2 (k=%d stride=%d pad=%d)", ksize, stride, pad);
    SSG_CHECK_ARG(x1 != nullptr || c1 == 0, "conv2d_fwd_tc: x1 missing");
    const int oh = (h + 2 * pad - ksize) / stride + 1, ow = (w + 2 * pad - ksize) / stride + 1;
    SSG_CHECK_ARG(oh > 0 && ow > 0, "conv2d_fwd_tc: empty output");
    TapClass c;
    memset(&c, 0, sizeof(c));
    c.ntaps = ksize * ksize;
    for (int r = 0; r < ksize; ++r)
        for (int q = 0; q < ksize; ++q) {
            const int t = r * ksize + q;
            c.dy[t] = (int8_t)(r - pad); c.dx[t] = (int8_t)(q - pad); c.wt[t] = (int8_t)t;
        }
    if (stride == 1 && 2 * pad == ksize - 1 && use_halo_kernel())
        return run_conv_halo(x0, c0, x1, c1, w_packed, ksize * ksize, bias, bias_n, y, n, h, w, cout, ksize, 0, act, slope, stats,
                             (cudaStream_t)s);
    SSG_CHECK_ARG(stats == nullptr, "conv2d_fwd_tc: fused statistics need a stride-1 same-size convolution (query ssg_conv2d_fwd_tc_has_stats)");
    return run_conv(x0, c0, x1, c1, w_packed, ksize * ksize, bias, bias_n, y, n, h, w, oh, ow, oh, ow, cout, stride, 1, &c, 1, act, slope,
                    (cudaStream_t)s);
}

int ssg_conv2d_fwd_tc_has_stats(int ksize, int stride, int pad) {
    return (stride == 1 && 2 * pad == ksize - 1 && (ksize == 1 || ksize == 3) && use_halo_kernel()) ? 1 : 0;
}

int ssg_conv2d_dgrad_tc_can_acc(int ksize, int stride, int pad) {
    return (stride == 1 && 2 * pad == ksize - 1 && (ksize == 1 || ksize == 3) && use_halo_kernel()) ? 1 : 0;
}

// dx += data gradient (same-size stride-1 convolutions only): the epilogue's TMA store becomes a TMA reduce-add, so the
// gradient contributions of two consumers of one activation (BasicBlock's conv1 + shortcut, archs.py:229-234; SPADE's
// x2map + modulation, normalization.py:112-120) meet in one buffer without a separate addition pass.
int ssg_conv2d_dgrad_tc_acc(const void* dy, const void* w_packed, void* dx, int n, int h, int w, int cin, int cout, int ksize, int stride,
                            int pad, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && cout % 8 == 0 && cin % 8 == 0,
                  "conv2d_dgrad_tc_acc: stored channel counts must be multiples of 8 (cin=%d cout=%d)", cin, cout);
    if (!ssg_conv2d_dgrad_tc_can_acc(ksize, stride, pad)) {
        set_error("conv2d_dgrad_tc_acc: only same-size stride-1 1x1 / 3x3 convolutions accumulate (k=%d stride=%d pad=%d)", ksize, stride, pad);
        return SSG_ERR_UNSUPPORTED;
    }
    return run_conv_halo(dy, cout, nullptr, 0, w_packed, ksize * ksize, nullptr, 0, dx, n, h, w, cin, ksize, 1, 0, 0.f, nullptr,
                         (cudaStream_t)s, 1);
}

// Data gradient of a convolution whose input was the VIRTUAL concatenation [x0 | x1] (archs.py:651-667): channels [0, c0) of the
// gradient are written (or added) to dx0, the rest to dx1 -- torch.cat's backward (the split copy) disappears.
int ssg_conv2d_dgrad_tc_split(const void* dy, const void* w_packed, void* dx0, int c0, void* dx1, int c1, int n, int h, int w, int cout,
                              int ksize, int stride, int pad, int accumulate, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && c0 > 0 && c1 > 0 && cout > 0 && cout % 8 == 0 && c0 % 64 == 0 && c1 % 8 == 0 && dx0 && dx1,
                  "conv2d_dgrad_tc_split: c0 %% 64 == 0, c1 %% 8 == 0, cout %% 8 == 0 required (c0=%d c1=%d cout=%d)", c0, c1, cout);
    if (!ssg_conv2d_dgrad_tc_can_acc(ksize, stride, pad)) {
        set_error("conv2d_dgrad_tc_split: only same-size stride-1 1x1 / 3x3 convolutions (k=%d stride=%d pad=%d)", ksize, stride, pad);
        return SSG_ERR_UNSUPPORTED;
    }
    return run_conv_halo(dy, cout, nullptr, 0, w_packed, ksize * ksize, nullptr, 0, dx0, n, h, w, c0 + c1, ksize, 1, 0, 0.f, nullptr,
                         (cudaStream_t)s, accumulate ? 1 : 0, dx1, c0);
}

static int dgrad_tc_impl(const void* dy, const void* w_packed, void* dx, int n, int h, int w, int cin, int cout, int ksize, int stride,
                         int pad, const void* mask, int mask_act, float mask_slope, ssg_stream_t s);

int ssg_conv2d_dgrad_tc(const void* dy, const void* w_packed, void* dx, int n, int h, int w, int cin, int cout, int ksize, int stride,
                        int pad, ssg_stream_t s) {
    return dgrad_tc_impl(dy, w_packed, dx, n, h, w, cin, cout, ksize, stride, pad, nullptr, 0, 0.f, s);
}

int ssg_conv2d_dgrad_tc_mask_supported(int ksize, int stride, int pad) { return (ksize == 3 && stride == 2) ? 1 : 0; }

int ssg_conv2d_dgrad_tc_mask(const void* dy, const void* w_packed, void* dx, const void* producer_out, int producer_act, float producer_slope,
                             int n, int h, int w, int cin, int cout, int ksize, int stride, int pad, ssg_stream_t s) {
    SSG_CHECK_ARG(producer_out != nullptr && ssg_conv2d_dgrad_tc_mask_supported(ksize, stride, pad),
                  "conv2d_dgrad_tc_mask: only the stride-2 3x3 data gradient carries the producer's activation backward (k=%d stride=%d)",
                  ksize, stride);
    return dgrad_tc_impl(dy, w_packed, dx, n, h, w, cin, cout, ksize, stride, pad, producer_out, producer_act, producer_slope, s);
}

static int dgrad_tc_impl(const void* dy, const void* w_packed, void* dx, int n, int h, int w, int cin, int cout, int ksize, int stride,
                         int pad, const void* mask, int mask_act, float mask_slope, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && cout % 8 == 0 && cin % 8 == 0,
                  "conv2d_dgrad_tc: stored channel counts must be multiples of 8 (cin=%d cout=%d)", cin, cout);
    SSG_CHECK_ARG((ksize == 1 || ksize == 3) && pad >= 0 && pad < ksize && (stride == 1 || (stride == 2 && ksize == 3)),
                  "conv2d_dgrad_tc: kernel 1 or 3 at stride 1, kernel 3 at stride 2 (k=%d stride=%d)", ksize, stride);
    const int oh = (h + 2 * pad - ksize) / stride + 1, ow = (w + 2 * pad - ksize) / stride + 1;
    SSG_CHECK_ARG(oh > 0 && ow > 0, "conv2d_dgrad_tc: empty dy");
    if (stride == 1 && 2 * pad == ksize - 1 && use_halo_kernel())
        // dx[i] = sum_r dy[i + pad - r] W[r]: halo origin i0 - pad, tap r reads halo row (k - 1 - r)
        return run_conv_halo(dy, cout, nullptr, 0, w_packed, ksize * ksize, nullptr, 0, dx, n, h, w, cin, ksize, 1, 0, 0.f, nullptr,
                             (cudaStream_t)s);
    if (stride == 2 && ksize == 3 && pad == 1 && use_halo_kernel() && dgrad_s2_halo_supported(h, w, cin, cout))
        return run_dgrad_s2_halo(dy, cout, w_packed, dx, n, h, w, cin, mask, mask_act, mask_slope, (cudaStream_t)s);
    TapClass c[4];
    memset(c, 0, sizeof(c));
    int ncls = 0;
    // dx[i] = sum_r dy[(i + pad - r) / stride] W[r] over the taps where the division is exact: one class per parity of i
    for (int py = 0; py < stride; ++py)
        for (int px = 0; px < stride; ++px) {
            TapClass& k = c[ncls++];
            k.add_y = (int8_t)py; k.add_x = (int8_t)px;
            for (int r = 0; r < ksize; ++r)
                for (int q = 0; q < ksize; ++q) {
                    const int ny = py + pad - r, nx = px + pad - q;
                    if (((ny % stride) + stride) % stride != 0 || ((nx % stride) + stride) % stride != 0) continue;
                    const int t = k.ntaps++;
                    k.dy[t] = (int8_t)(ny >= 0 ? ny / stride : -((-ny) / stride));
                    k.dx[t] = (int8_t)(nx >= 0 ? nx / stride : -((-nx) / stride));
                    k.wt[t] = (int8_t)(r * ksize + q);
                }
            SSG_CHECK_ARG(k.ntaps > 0, "conv2d_dgrad_tc: empty parity class");
        }
    const int th_ = (h + stride - 1) / stride, tw_ = (w + stride - 1) / stride;
    return run_conv(dy, cout, nullptr, 0, w_packed, ksize * ksize, nullptr, 0, dx, n, oh, ow, th_, tw_, h, w, cin, 1, stride, c, ncls, 0, 0.f,
                    (cudaStream_t)s, mask, mask_act, mask_slope);
}

}  // extern "C"

// =====================================================================================================
// Weight gradient: dW[co][ci][r][s] = sum over pixels of dy[pix][co] * x[pix + (r,s) - pad][ci]
// GEMM view per (tap, 64-channel chunk) "unit": D[M = 128 rows = 2 units x 64 ci][N = co tile] += A^T B with the
// pixel index as the reduction dimension, i.e. BOTH operands are MN-major: the same {64 ch x 128 px} TMA boxes as
// the forward kernel, read by tcgen05.mma with a_major = b_major = MN (no transposes anywhere).
// A CTA owns G accumulators (G * N <= 512 TMEM columns) for one co tile and walks a strided subset of the pixel
// tiles (split-K); partial sums are combined with fp32 red.global.add into the OIHW gradient.
// =====================================================================================================
namespace ssg {
namespace tc {

struct WgradParams {
    float* dw;                  // OIHW fp32, pre-zeroed
    int N, H, W;                // dy spatial dims (== x dims, stride 1)
    int cout, cin;              // REAL channel counts: extents of dw (stored tensors may be channel-padded)
    int tw_log2, th_log2, tiles_x, tiles_y, m_tiles;
    int taps, kw, pad;
    int in_mul;                 // convolution stride: x pixel = in_mul * dy pixel + tap - pad
    int chunks0, chunks1;
    int units;                  // taps * (chunks0 + chunks1)
};

template <int BN>   // co tile: 64 or 128
struct WgradCfg {
    static constexpr int G = 512 / BN;               // accumulators per CTA (unit pairs)
    static constexpr int NJ = BN / 64;               // dy boxes per pixel tile
    static constexpr int XS = 4;                     // x ring slots (each = one unit pair = 2 boxes)
    static constexpr int BOX = BM * BK * 2;          // 16 KB
    static constexpr int DY_BYTES = NJ * BOX;
    static constexpr int X_OFFSET = 2 * DY_BYTES;
    static constexpr int BAR_OFFSET = X_OFFSET + XS * 2 * BOX;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS) conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmX0,
                                                                     const __grid_constant__ CUtensorMap tmX1,
                                                                     const __grid_constant__ CUtensorMap tmDY, const WgradParams p) {
    using C = WgradCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* dy_full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFFSET);
    uint64_t* dy_empty = dy_full + 2;
    uint64_t* x_full = dy_empty + 2;
    uint64_t* x_empty = x_full + C::XS;
    uint64_t* acc_full = x_empty + C::XS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TW = 1 << p.tw_log2, TH = 1 << p.th_log2;
    const int chunks = p.chunks0 + p.chunks1;
    const int unit0 = blockIdx.x * 2 * C::G;
    int n_units = p.units - unit0;
    if (n_units > 2 * C::G) n_units = 2 * C::G;
    const int pairs = (n_units + 1) >> 1;
    const int co0 = blockIdx.y * BN;
    const int n_iter = (p.m_tiles - (int)blockIdx.z + (int)gridDim.z - 1) / (int)gridDim.z;   // pixel tiles of this CTA

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&dy_full[i], 1); mbar_init(&dy_empty[i], 1); }
        for (int i = 0; i < C::XS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (lane == 0) {
            int xc = 0;
            for (int it = 0; it < n_iter; ++it) {
                int tile = blockIdx.z + it * gridDim.z;
                const int tx = tile % p.tiles_x; tile /= p.tiles_x;
                const int ty = tile % p.tiles_y;
                const int tn = tile / p.tiles_y;
                const int w0 = tx * TW, h0 = ty * TH, img0 = tn * (BM >> (p.tw_log2 + p.th_log2));
                const int b = it & 1;
                mbar_wait(&dy_empty[b], ((it >> 1) & 1) ^ 1);
                mbar_expect_tx(&dy_full[b], C::DY_BYTES);
                for (int j = 0; j < C::NJ; ++j)
                    tma_load_4d(smem + b * C::DY_BYTES + j * C::BOX, &tmDY, co0 + 64 * j, w0, h0, img0, &dy_full[b]);
                for (int g = 0; g < pairs; ++g, ++xc) {
                    const int slot = xc % C::XS;
                    mbar_wait(&x_empty[slot], ((xc / C::XS) & 1) ^ 1);
                    mbar_expect_tx(&x_full[slot], 2 * C::BOX);
                    for (int e = 0; e < 2; ++e) {
                        int u = unit0 + 2 * g + e;
                        if (u >= p.units) u = unit0;                       // dummy half: rows are never stored
                        const int tap = u / chunks, ch = u - tap * chunks;
                        const int r = tap / p.kw, s = tap - r * p.kw;
                        uint8_t* dst = smem + C::X_OFFSET + (slot * 2 + e) * C::BOX;
                        const int cx = p.in_mul * w0 + s - p.pad, cy = p.in_mul * h0 + r - p.pad;
                        if (ch < p.chunks0) tma_load_4d(dst, &tmX0, ch * BK, cx, cy, img0, &x_full[slot]);
                        else tma_load_4d(dst, &tmX1, (ch - p.chunks0) * BK, cx, cy, img0, &x_full[slot]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // whole warp walks the (uniform) loop; the elected lane issues (see conv_tc_fwd_kernel)
        constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 1, 1);     // both operands MN-major
        const bool leader = elect_one();
        // MN-major SW128: LBO = distance between 64-element MN atoms (one box), SBO = 8 pixel rows
        const uint32_t hi = desc_hi(1024, 2);
        int xc = 0;
        for (int it = 0; it < n_iter; ++it) {
            const int b = it & 1;
            mbar_wait(&dy_full[b], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t dy_lo = desc_lo(smem_u32(smem + b * C::DY_BYTES), C::BOX);
            for (int g = 0; g < pairs; ++g, ++xc) {
                const int slot = xc % C::XS;
                mbar_wait(&x_full[slot], (xc / C::XS) & 1);
                tc_fence_after();
                const uint32_t x_lo = desc_lo(smem_u32(smem + C::X_OFFSET + slot * 2 * C::BOX), C::BOX);
                if (leader) {
                    umma_bf16_lohi(tmem_base + (uint32_t)(g * BN), x_lo, hi, dy_lo, hi, idesc, (uint32_t)it);
#pragma unroll
                    for (int k = 1; k < BM / 16; ++k)    // 16 pixels per MMA: +2048 bytes (>>4 == 128)
                        umma_bf16_lohi(tmem_base + (uint32_t)(g * BN), x_lo + (uint32_t)(128 * k), hi, dy_lo + (uint32_t)(128 * k), hi, idesc, 1u);
                    umma_commit(&x_empty[slot]);
                }
                __syncwarp();
            }
            if (leader) umma_commit(&dy_empty[b]);
            __syncwarp();
        }
        if (leader) umma_commit(acc_full);
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;                  // accumulator row: unit (row >> 6), ci (row & 63)
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int kk = p.taps;                           // kh * kw
        for (int g = 0; g < pairs; ++g) {
            const int u = unit0 + 2 * g + (row >> 6);
            const int tap = u / chunks, ch = u - tap * chunks;
            const int ci = ch * BK + (row & 63);
            const bool row_ok = u < p.units && n_iter > 0 && ci < p.cin;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BN + c0), v);
                tmem_ld_wait();
                if (!row_ok) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int co = co0 + c0 + j;
                    if (co < p.cout) atomicAdd(p.dw + ((long long)co * p.cin + ci) * kk + tap, __uint_as_float(v[j]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int BN>
static int launch_wgrad(const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap& dy, const WgradParams& p, cudaStream_t st) {
    using C = WgradCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        SSG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
        attr_set = true;
    }
    const int groups = (p.units + 2 * C::G - 1) / (2 * C::G);
    const int co_tiles = (p.cout + BN - 1) / BN;
    // Split-K factor: one CTA per SM is resident (512 TMEM columns each) and every CTA ends with a fixed-size flush of its
    // accumulators into L2, so pick the factor that minimises waves * (tiles per CTA + flush) -- see run_wgrad_halo
    const int pairs = groups * co_tiles, sms = sm_count_cached();
    int splits = 1;
    long long best = -1;
    for (int zc = 1; zc <= (2 * sms + pairs - 1) / pairs + 1 && zc <= p.m_tiles; ++zc) {
        const long long waves = ((long long)pairs * zc + sms - 1) / sms;
        const long long cost = waves * ((p.m_tiles + zc - 1) / zc + 24);
        if (best < 0 || cost < best) { best = cost; splits = zc; }
    }
    dim3 grid((unsigned)groups, (unsigned)co_tiles, (unsigned)splits);
    conv_tc_wgrad_kernel<BN><<<grid, NUM_THREADS, C::TOTAL, st>>>(x0, x1, dy, p);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // namespace tc
}  // namespace ssg

static int wgrad_tc_impl(const void* x0, int c0, const void* x1, int c1, const void* dy, int cout, float* dw_oihw, int cout_real,
                         int cin_real, int n, int h, int w, int ksize, int stride, int pad, bool accumulate, ssg_stream_t s);

extern "C" int ssg_conv2d_wgrad_tc(const void* x0, int c0, const void* x1, int c1, const void* dy, int cout, float* dw_oihw,
                                   int cout_real, int cin_real, int n, int h, int w, int ksize, int stride, int pad, ssg_stream_t s) {
    return wgrad_tc_impl(x0, c0, x1, c1, dy, cout, dw_oihw, cout_real, cin_real, n, h, w, ksize, stride, pad, false, s);
}

extern "C" int ssg_conv2d_wgrad_tc_acc(const void* x0, int c0, const void* x1, int c1, const void* dy, int cout, float* dw_oihw,
                                       int cout_real, int cin_real, int n, int h, int w, int ksize, int stride, int pad, ssg_stream_t s) {
    return wgrad_tc_impl(x0, c0, x1, c1, dy, cout, dw_oihw, cout_real, cin_real, n, h, w, ksize, stride, pad, true, s);
}

static int wgrad_tc_impl(const void* x0, int c0, const void* x1, int c1, const void* dy, int cout, float* dw_oihw, int cout_real,
                         int cin_real, int n, int h, int w, int ksize, int stride, int pad, bool accumulate, ssg_stream_t s) {
    using namespace ssg;
    using namespace ssg::tc;
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && cout > 0 && cout % 8 == 0 && c0 > 0 && c0 % 8 == 0 && c1 >= 0 && c1 % 8 == 0 &&
                      (c1 == 0 || c0 % 64 == 0) && cout_real > 0 && cout_real <= cout && cin_real > 0 && cin_real <= c0 + c1,
                  "conv2d_wgrad_tc: stored channel counts must be multiples of 8 (c0=%d c1=%d cout=%d)", c0, c1, cout);
    SSG_CHECK_ARG((ksize == 1 || ksize == 3) && pad >= 0 && pad < ksize && (stride == 1 || stride == 2),
                  "conv2d_wgrad_tc: kernel 1 or 3, stride 1 or 2");
    const int oh = (h + 2 * pad - ksize) / stride + 1, ow = (w + 2 * pad - ksize) / stride + 1;
    SSG_CHECK_ARG(oh > 0 && ow > 0, "conv2d_wgrad_tc: empty dy");
    const int taps = ksize * ksize;
    cudaStream_t st = (cudaStream_t)s;
    if (!accumulate) SSG_CHECK_CUDA(cudaMemsetAsync(dw_oihw, 0, sizeof(float) * (size_t)cout_real * cin_real * taps, st));
    static const bool no_thin = getenv("SSG_NO_THIN_WGRAD") != nullptr;
    if (ksize == 3 && stride == 1 && pad == 1 && c0 == 8 && c1 == 0 && !no_thin)      // image stems, SPADE's 8-channel maps
        return run_wgrad_thin(x0, dy, cout, dw_oihw, cout_real, cin_real, n, h, w, st);
    if (ksize == 3 && stride == 1 && pad == 1 && use_halo_kernel())
        return run_wgrad_halo(x0, c0, x1, c1, dy, cout, dw_oihw, cout_real, cin_real, n, h, w, st);
    int twl, thl;
    pick_tile(oh, ow, twl, thl);          // pixel tiles enumerate dy; x is gathered at stride * pixel + tap - pad
    const int TW = 1 << twl, TH = 1 << thl, NB = BM / (TW * TH);
    WgradParams p;
    p.dw = dw_oihw; p.N = n; p.H = oh; p.W = ow; p.cout = cout_real; p.cin = cin_real; p.in_mul = stride;
    p.tw_log2 = twl; p.th_log2 = thl; p.tiles_x = (ow + TW - 1) / TW; p.tiles_y = (oh + TH - 1) / TH;
    p.m_tiles = p.tiles_x * p.tiles_y * ((n + NB - 1) / NB);
    p.taps = taps; p.kw = ksize; p.pad = pad; p.chunks0 = (c0 + 63) / 64; p.chunks1 = (c1 + 63) / 64; p.units = taps * (p.chunks0 + p.chunks1);
    CUtensorMap mx0, mx1, mdy;
    int rc = encode_act_map(&mx0, x0, c0, n, h, w, twl, thl, stride);
    if (rc) return rc;
    mx1 = mx0;
    if (c1 > 0) {
        rc = encode_act_map(&mx1, x1, c1, n, h, w, twl, thl, stride);
        if (rc) return rc;
    }
    rc = encode_act_map(&mdy, dy, cout, n, oh, ow, twl, thl, 1);
    if (rc) return rc;
    if (cout_real > 64) return launch_wgrad<128>(mx0, mx1, mdy, p, st);
    return launch_wgrad<64>(mx0, mx1, mdy, p, st);
}
