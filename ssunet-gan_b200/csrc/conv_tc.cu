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

namespace ssg {
namespace tc {

constexpr int BM = 128;          // pixels per tile == UMMA M
constexpr int BK = 64;           // bf16 channels per k-block (128 bytes == swizzle span)
constexpr int A_BYTES = BM * BK * 2;
constexpr int NUM_THREADS = 192;

struct ConvTcParams {
    bf16* y;
    const float* bias;
    int N, H, W;                 // output (== input) batch / spatial dims
    int cout;                    // real output channels; row stride of y
    int tw_log2, th_log2;        // tile = TW x TH x NB pixels (product 128)
    int tiles_x, tiles_y;
    int taps, kw, pad;
    int chunks0, chunks1;        // 64-channel chunks taken from x0 / x1
    int act;
    float slope;
};

template <int BN>
struct SmemLayout {
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN >= 128) ? 3 : 4;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // + barriers + alignment slack
    static constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS) conv_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                                   const __grid_constant__ CUtensorMap tmA1,
                                                                   const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
    using L = SmemLayout<BN>;
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
    const int num_kb = p.taps * chunks;

    if (threadIdx.x == 0) {
        for (int i = 0; i < L::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, L::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tmA0);
            tma_prefetch_desc(&tmB);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int stage = kb % L::STAGES;
                const uint32_t phase = (kb / L::STAGES) & 1;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                const int tap = kb / chunks, ch = kb - tap * chunks;
                const int r = tap / p.kw, s = tap - r * p.kw;
                uint8_t* sa = smem + stage * L::STAGE_BYTES;
                uint8_t* sb = sa + A_BYTES;
                if (ch < p.chunks0) tma_load_4d(sa, &tmA0, ch * BK, w0 + s - p.pad, h0 + r - p.pad, img0, &full_bar[stage]);
                else tma_load_4d(sa, &tmA1, (ch - p.chunks0) * BK, w0 + s - p.pad, h0 + r - p.pad, img0, &full_bar[stage]);
                tma_load_3d(sb, &tmB, ch * BK, n0, tap, &full_bar[stage]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int stage = kb % L::STAGES;
                const uint32_t phase = (kb / L::STAGES) & 1;
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
                const uint64_t da = make_desc_kmajor_sw128(sa);
                const uint64_t db = make_desc_kmajor_sw128(sa + A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)   // +32 bytes (>>4 == 2) per 16-element K step inside the swizzle atom
                    umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                umma_commit(&empty_bar[stage]);     // frees the smem slot once these MMAs have read it
            }
            umma_commit(tmem_full_bar);             // accumulator complete
        }
    } else {
        // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) ----
        const int q = warp & 3;
        const int m = q * 32 + lane;                // accumulator row == pixel index inside the tile
        const int twi = m & (TW - 1), thi = (m >> p.tw_log2) & (TH - 1), nbi = m >> (p.tw_log2 + p.th_log2);
        const int ox = w0 + twi, oy = h0 + thi, on = img0 + nbi;
        const bool row_ok = ox < p.W && oy < p.H && on < p.N;
        bf16* yrow = p.y + (((long long)on * p.H + oy) * p.W + ox) * p.cout + n0;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const bool vec_ok = (p.cout % 8 == 0);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
            if (!row_ok) continue;
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = n0 + c0 + j;
                float t = __uint_as_float(v[j]);
                if (p.bias != nullptr && n < p.cout) t += p.bias[n];
                f[j] = apply_act(t, p.act, p.slope);
            }
            if (vec_ok && n0 + c0 + 16 <= p.cout) {
                Vec<bf16> o;
                o.set(f); o.store(yrow + c0);
                o.set(f + 8); o.store(yrow + c0 + 8);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (n0 + c0 + j < p.cout) yrow[c0 + j] = __float2bfloat16_rn(f[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, L::TMEM_COLS);
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
int encode_bf16_map(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, CUtensorMapSwizzle swizzle) {
    auto enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return SSG_ERR_CUDA; }
    uint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return SSG_ERR_CUDA; }
    return SSG_OK;
}

static int ilog2_ceil(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

template <int BN>
static int launch_fwd(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvTcParams& p, int m_tiles,
                      cudaStream_t st) {
    using L = SmemLayout<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        SSG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_fwd_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    dim3 grid((unsigned)m_tiles, (unsigned)((p.cout + BN - 1) / BN));
    conv_tc_fwd_kernel<BN><<<grid, NUM_THREADS, L::TOTAL, st>>>(a0, a1, b, p);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // namespace tc
}  // namespace ssg
using namespace ssg;
using namespace ssg::tc;

extern "C" {

int ssg_conv2d_fwd_tc(const void* x0, int c0, const void* x1, int c1, const void* w_packed, const float* bias, void* y, int n, int h,
                      int w, int cout, int ksize, int pad, int act, float slope, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && cout > 0 && c0 > 0 && c0 % 64 == 0 && c1 >= 0 && c1 % 64 == 0,
                  "conv2d_fwd_tc: channels must be multiples of 64 (c0=%d c1=%d)", c0, c1);
    SSG_CHECK_ARG((ksize == 1 || ksize == 3) && pad >= 0 && 2 * pad == ksize - 1, "conv2d_fwd_tc: only 1x1/p0 and 3x3/p1 (same-size) convolutions");
    SSG_CHECK_ARG(x1 != nullptr || c1 == 0, "conv2d_fwd_tc: x1 missing");
    const int cin = c0 + c1, taps = ksize * ksize;
    int twl = ilog2_ceil(w); if (twl > 7) twl = 7;
    int thl = ilog2_ceil(h); if (thl > 7 - twl) thl = 7 - twl;
    const int TW = 1 << twl, TH = 1 << thl, NB = BM / (TW * TH);
    ConvTcParams p;
    p.y = (bf16*)y; p.bias = bias; p.N = n; p.H = h; p.W = w; p.cout = cout;
    p.tw_log2 = twl; p.th_log2 = thl;
    p.tiles_x = (w + TW - 1) / TW; p.tiles_y = (h + TH - 1) / TH;
    const int tiles_n = (n + NB - 1) / NB;
    p.taps = taps; p.kw = ksize; p.pad = pad; p.chunks0 = c0 / 64; p.chunks1 = c1 / 64; p.act = act; p.slope = slope;
    CUtensorMap ma0, ma1, mb;
    {
        uint64_t dims[4] = {(uint64_t)c0, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        uint64_t str[3] = {(uint64_t)c0 * 2, (uint64_t)w * c0 * 2, (uint64_t)h * w * c0 * 2};
        uint32_t box[4] = {64, (uint32_t)TW, (uint32_t)TH, (uint32_t)NB};
        int rc = encode_bf16_map(&ma0, x0, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        ma1 = ma0;
        if (c1 > 0) {
            uint64_t dims1[4] = {(uint64_t)c1, (uint64_t)w, (uint64_t)h, (uint64_t)n};
            uint64_t str1[3] = {(uint64_t)c1 * 2, (uint64_t)w * c1 * 2, (uint64_t)h * w * c1 * 2};
            rc = encode_bf16_map(&ma1, x1, 4, dims1, str1, box, CU_TENSOR_MAP_SWIZZLE_128B);
            if (rc) return rc;
        }
    }
    const int BN = cout >= 128 ? 128 : (cout > 16 ? 64 : 16);
    {
        uint64_t dims[3] = {(uint64_t)cin, (uint64_t)cout, (uint64_t)taps};
        uint64_t str[2] = {(uint64_t)cin * 2, (uint64_t)cout * cin * 2};
        uint32_t box[3] = {64, (uint32_t)BN, 1};
        int rc = encode_bf16_map(&mb, w_packed, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const int m_tiles = p.tiles_x * p.tiles_y * tiles_n;
    cudaStream_t st = (cudaStream_t)s;
    if (BN == 128) return launch_fwd<128>(ma0, ma1, mb, p, m_tiles, st);
    if (BN == 64) return launch_fwd<64>(ma0, ma1, mb, p, m_tiles, st);
    return launch_fwd<16>(ma0, ma1, mb, p, m_tiles, st);
}

}  // extern "C"
