// Self-conditioned SPADE forward (normalization.py:106-122 with segmap = x) as ONE kernel per 16 x 16 output tile:
//     seg  = x2map(x)                 C -> label_nc (<= 8), 3x3, bias          (stored with 8 channels, bf16)
//     actv = relu(mlp_shared(seg))    label_nc -> h (<= 8), 3x3, bias          (stored with 8 channels, bf16)
//     g|b  = [mlp_gamma | mlp_beta](actv)   h -> 2C, 3x3, bias                 (bf16, optional output: the backward reads it)
//     y    = x * (1 + g) + b
// The unfused chain moves ~7 C H W elements through HBM (x read twice, g|b written and read) and its three convolutions are too
// thin for a 128-row tcgen05 tile (N = 8, K = 27 .. 72).  Here x is staged ONCE per tile with a 3-pixel
// halo (22 x 22 x C bf16), seg is computed over the 20 x 20 halo-2 region, actv over 18 x 18, g|b over the 16 x 16 tile, all
// three contractions as warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate; N = 8 is its native width) with the operands
// gathered straight from shared memory by ldmatrix (one 8-channel pixel = one 16-byte matrix row, two taps per k-step), and the
// modulation is applied to the accumulators in registers.  Every intermediate is rounded to bf16 exactly where the unfused chain
// stores it, and seg / actv outside the image are forced to zero (the zero padding the unfused convolutions see), so the
// results agree with the chain to fp32 accumulation order.
// STATUS: opt-in (SSG_SPADE_FUSED=1 / ops.set_spade_fused): built and unit-tested against the unfused chain before it becomes
// the default.  Supported: C in {64, 128} (levels 0 and 1 of the U-Net: ~80 % of the SPADE time), label_nc <= 8, h <= 8.
#include "common.cuh"

namespace ssg {
namespace spf {

constexpr int TILE = 16, XR = TILE + 6, SR = TILE + 4, AR = TILE + 2;      // x / seg / actv region edge: 22 / 20 / 18
constexpr int K3P = 88;                                                    // padded K pitch (bf16) of the 80-wide operands

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

struct Params {
    const bf16* x;        // [N, H, W, C]
    const bf16* w1;       // [9][8][C]      x2map, rows n >= label_nc are zero
    const bf16* w2;       // [8][80]        mlp_shared, k = tap * 8 + ci (k >= 72 zero), rows n >= h zero
    const bf16* w3;       // [2C][80]       gamma rows then beta rows, k = tap * 8 + ci
    const float* b1;      // [8]
    const float* b2;      // [8]
    const float* b3;      // [2C]
    bf16* seg;            // [N, H, W, 8]
    bf16* actv;           // [N, H, W, 8]
    bf16* gb;             // [N, H, W, 2C] or null
    bf16* y;              // [N, H, W, C]
    int N, H, W;
};

template <int C>
struct Smem {
    static constexpr int W1P = C + 8;                                  // padded row pitch (bf16) of the x2map operand
    static constexpr int XS = 0;                                       // [XR*XR][C] bf16, 16-byte chunks XOR-swizzled by (pixel & 7)
    static constexpr int SEG = XS + XR * XR * C * 2;                   // [SR*SR + 1][8] bf16 (last row: zeros)
    static constexpr int ACT = SEG + (SR * SR + 1) * 16;               // [AR*AR + 1][8] bf16 (last row: zeros)
    static constexpr int W1 = ACT + (AR * AR + 1) * 16;                // [9][8][W1P]
    static constexpr int W2 = W1 + 9 * 8 * W1P * 2;                    // [8][K3P]
    static constexpr int W3 = W2 + 8 * K3P * 2;                        // [2C][K3P]
    static constexpr int TOTAL = W3 + 2 * C * K3P * 2;
};

template <int C>
__global__ void __launch_bounds__(256, 1) spade_fused_fwd_kernel(const Params p) {
    using S = Smem<C>;
    extern __shared__ __align__(128) uint8_t smem[];
    bf16* xs = reinterpret_cast<bf16*>(smem + S::XS);
    bf16* segs = reinterpret_cast<bf16*>(smem + S::SEG);
    bf16* acts = reinterpret_cast<bf16*>(smem + S::ACT);
    bf16* w1s = reinterpret_cast<bf16*>(smem + S::W1);
    bf16* w2s = reinterpret_cast<bf16*>(smem + S::W2);
    bf16* w3s = reinterpret_cast<bf16*>(smem + S::W3);
    constexpr int CH = C / 8;                                          // 16-byte chunks per pixel of x
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int img = blockIdx.z, ty0 = blockIdx.y * TILE, tx0 = blockIdx.x * TILE;
    const bf16* ximg = p.x + (long long)img * p.H * p.W * C;

    // ---- stage 0: x tile with halo 3 (zeros outside the image), the three weight operands, the zero rows ----
    for (int i = tid; i < XR * XR * CH; i += 256) {
        const int q = i / CH, k = i - q * CH;
        const int iy = ty0 - 3 + q / XR, ix = tx0 - 3 + q % XR;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) v = __ldg(reinterpret_cast<const uint4*>(ximg + ((long long)iy * p.W + ix) * C) + k);
        *reinterpret_cast<uint4*>(smem + S::XS + (size_t)q * C * 2 + ((k ^ (q & 7)) << 4)) = v;
    }
    for (int i = tid; i < 9 * 8 * CH; i += 256) {                       // w1 [72 rows][C] -> pitch W1P
        const int row = i / CH, k = i - row * CH;
        *reinterpret_cast<uint4*>(w1s + (size_t)row * S::W1P + k * 8) = __ldg(reinterpret_cast<const uint4*>(p.w1 + (size_t)row * C) + k);
    }
    for (int i = tid; i < (8 + 2 * C) * 10; i += 256) {                 // w2 [8][80] and w3 [2C][80] -> pitch K3P
        const int row = i / 10, k = i - row * 10;
        const bf16* src = row < 8 ? p.w2 + (size_t)row * 80 : p.w3 + (size_t)(row - 8) * 80;
        bf16* dst = row < 8 ? w2s + (size_t)row * K3P : w3s + (size_t)(row - 8) * K3P;
        *reinterpret_cast<uint4*>(dst + k * 8) = __ldg(reinterpret_cast<const uint4*>(src) + k);
    }
    if (tid < 2) *reinterpret_cast<uint4*>(tid == 0 ? segs + SR * SR * 8 : acts + AR * AR * 8) = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();

    const uint32_t xs_a = smem_addr(xs), seg_a = smem_addr(segs), act_a = smem_addr(acts);
    const uint32_t w1_a = smem_addr(w1s), w2_a = smem_addr(w2s), w3_a = smem_addr(w3s);
    const int arow = lane & 15, ahalf = lane >> 4;                       // ldmatrix.x4 address roles: tile row, k half
    const int bn = lane & 7, bhalf = (lane >> 3) & 1;                    // ldmatrix.x2 address roles: n row, k half

    // ---- stage A: seg = x2map(x) over the 20 x 20 region (25 M-tiles of 16 pixels), K = 9 taps x C ----
    for (int mt = warp; mt < (SR * SR) / 16; mt += 8) {
        const int pa = mt * 16 + arow;                                   // the pixel whose row address this lane supplies
        const int say = pa / SR, sax = pa - say * SR;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int q = (say + tap / 3) * XR + sax + tap % 3;          // xs pixel of this tap
            const uint32_t arow_a = xs_a + (uint32_t)q * (C * 2);
            const uint32_t brow_a = w1_a + (uint32_t)((tap * 8 + bn) * S::W1P + bhalf * 8) * 2;
#pragma unroll
            for (int kc = 0; kc < C / 16; ++kc) {
                uint32_t a[4], b[2];
                ldsm_x4(a, arow_a + ((uint32_t)((2 * kc + ahalf) ^ (q & 7)) << 4));
                ldsm_x2(b, brow_a + kc * 32);
                mma_bf16(acc, a, b);
            }
        }
        // accumulator rows: pixels mt*16 + g and + g + 8; columns 2*t4, 2*t4 + 1
        const float bia0 = __ldg(p.b1 + 2 * t4), bia1 = __ldg(p.b1 + 2 * t4 + 1);
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int pp = mt * 16 + g + 8 * hrow;
            const int sy = pp / SR, sx = pp - sy * SR;
            const int iy = ty0 - 2 + sy, ix = tx0 - 2 + sx;
            const bool inside = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
            const uint32_t v = inside ? pack_bf16(acc[2 * hrow] + bia0, acc[2 * hrow + 1] + bia1) : 0u;
            *reinterpret_cast<uint32_t*>(segs + pp * 8 + 2 * t4) = v;
            if (inside && sy >= 2 && sy < 2 + TILE && sx >= 2 && sx < 2 + TILE)
                *reinterpret_cast<uint32_t*>(p.seg + (((long long)img * p.H + iy) * p.W + ix) * 8 + 2 * t4) = v;
        }
    }
    __syncthreads();

    // ---- stage B: actv = relu(mlp_shared(seg)) over the 18 x 18 region (21 M-tiles, the last one ragged), K = 9 taps x 8 ----
    for (int mt = warp; mt < (AR * AR + 15) / 16; mt += 8) {
        const int pa = mt * 16 + arow;
        const int aay = pa / AR, aax = pa - aay * AR;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int tap = 2 * j + ahalf;
            const int q = (pa < AR * AR && tap < 9) ? (aay + tap / 3) * SR + aax + tap % 3 : SR * SR;      // zero row otherwise
            uint32_t a[4], b[2];
            ldsm_x4(a, seg_a + (uint32_t)q * 16);
            ldsm_x2(b, w2_a + (uint32_t)(bn * K3P + j * 16 + bhalf * 8) * 2);
            mma_bf16(acc, a, b);
        }
        const float bia0 = __ldg(p.b2 + 2 * t4), bia1 = __ldg(p.b2 + 2 * t4 + 1);
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int pp = mt * 16 + g + 8 * hrow;
            if (pp >= AR * AR) continue;
            const int ay = pp / AR, ax = pp - ay * AR;
            const int iy = ty0 - 1 + ay, ix = tx0 - 1 + ax;
            const bool inside = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
            const uint32_t v = inside ? pack_bf16(fmaxf(acc[2 * hrow] + bia0, 0.f), fmaxf(acc[2 * hrow + 1] + bia1, 0.f)) : 0u;
            *reinterpret_cast<uint32_t*>(acts + pp * 8 + 2 * t4) = v;
            if (inside && ay >= 1 && ay < 1 + TILE && ax >= 1 && ax < 1 + TILE)
                *reinterpret_cast<uint32_t*>(p.actv + (((long long)img * p.H + iy) * p.W + ix) * 8 + 2 * t4) = v;
        }
    }
    __syncthreads();

    // ---- stage C: g|b over the 16 x 16 tile (16 M-tiles, two per warp) and the modulation in registers ----
    for (int mt = warp; mt < (TILE * TILE) / 16; mt += 8) {
        const int pa = mt * 16 + arow;
        const int oay = pa / TILE, oax = pa - oay * TILE;
        uint32_t a[5][4];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int tap = 2 * j + ahalf;
            const int q = tap < 9 ? (oay + tap / 3) * AR + oax + tap % 3 : AR * AR;
            ldsm_x4(a[j], act_a + (uint32_t)q * 16);
        }
        int oy[2], ox[2];
        bool ok[2];
        long long gpix[2];
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int pp = mt * 16 + g + 8 * hrow;
            oy[hrow] = pp / TILE; ox[hrow] = pp - oy[hrow] * TILE;
            const int iy = ty0 + oy[hrow], ix = tx0 + ox[hrow];
            ok[hrow] = iy < p.H && ix < p.W;
            gpix[hrow] = ((long long)img * p.H + iy) * p.W + ix;
        }
#pragma unroll 2
        for (int cj = 0; cj < C / 8; ++cj) {                             // 8 gamma channels and the 8 matching beta channels
            float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                uint32_t bgm[2], bbt[2];
                ldsm_x2(bgm, w3_a + (uint32_t)((cj * 8 + bn) * K3P + j * 16 + bhalf * 8) * 2);
                ldsm_x2(bbt, w3_a + (uint32_t)((C + cj * 8 + bn) * K3P + j * 16 + bhalf * 8) * 2);
                mma_bf16(ag, a[j], bgm);
                mma_bf16(ab, a[j], bbt);
            }
            const int ch = cj * 8 + 2 * t4;
            const float g0 = __ldg(p.b3 + ch), g1 = __ldg(p.b3 + ch + 1), e0 = __ldg(p.b3 + C + ch), e1 = __ldg(p.b3 + C + ch + 1);
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                if (!ok[hrow]) continue;
                // gamma / beta rounded to bf16 as the unfused chain stores them; x from the staged tile (centre region)
                const float gm0 = bf16_round(ag[2 * hrow] + g0), gm1 = bf16_round(ag[2 * hrow + 1] + g1);
                const float bt0 = bf16_round(ab[2 * hrow] + e0), bt1 = bf16_round(ab[2 * hrow + 1] + e1);
                const int q = (oy[hrow] + 3) * XR + ox[hrow] + 3;
                const uint32_t xv = *reinterpret_cast<const uint32_t*>(smem + S::XS + (size_t)q * C * 2 + ((cj ^ (q & 7)) << 4) + t4 * 4);
                const float x0 = __uint_as_float(xv << 16), x1 = __uint_as_float(xv & 0xffff0000u);
                const float y0 = fmaf(x0, 1.f + gm0, bt0), y1 = fmaf(x1, 1.f + gm1, bt1);
                *reinterpret_cast<uint32_t*>(p.y + gpix[hrow] * C + ch) = pack_bf16(y0, y1);
                if (p.gb) {
                    *reinterpret_cast<uint32_t*>(p.gb + gpix[hrow] * (2 * C) + ch) = pack_bf16(gm0, gm1);
                    *reinterpret_cast<uint32_t*>(p.gb + gpix[hrow] * (2 * C) + C + ch) = pack_bf16(bt0, bt1);
                }
            }
        }
    }
}

template <int C>
static int launch(const Params& p, cudaStream_t st) {
    using S = Smem<C>;
    static bool attr = false;
    if (!attr) {
        SSG_CHECK_CUDA(cudaFuncSetAttribute(spade_fused_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        attr = true;
    }
    dim3 grid((unsigned)((p.W + TILE - 1) / TILE), (unsigned)((p.H + TILE - 1) / TILE), (unsigned)p.N);
    spade_fused_fwd_kernel<C><<<grid, 256, S::TOTAL, st>>>(p);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}


// =====================================================================================================================
// Version 2 (C = 64, label_nc <= 3; opt-in with SSG_SPADE_FUSED=2, NOT yet run on a GPU when this was written).
// v1 measured 1.40 ms against 1.245 ms for the unfused chain at level 0; static analysis (profiles/r01_spade_fused_v1.txt)
// blames the shared-memory port: every one of v1's 2 285 MMAs per tile fetches both operands from shared memory, stage A
// alone re-reads each x element nine times (once per tap).  v2 changes three things:
//   1. persistent CTAs (grid = SM count): the mlp_shared / gamma|beta operands are loaded into shared memory ONCE per CTA
//      and the x2map operand lives in registers for the whole kernel;
//   2. stage A is turned around ("weights stationary"): P[tap * L + c][pixel] = sum_k W1[tap, c, k] x[pixel, k] is computed
//      once per x pixel with the WEIGHTS as the 16-row A operand (32 rows = 9 taps x L <= 3 classes, zero padded) and 8
//      pixels as the N dimension -- each x element is read from shared memory once -- and seg[pixel][c] is then the sum
//      of nine shifted P values (fp32, through a 62 KB shared buffer);
//   3. stage C keeps the activation fragments of BOTH of a warp's M-tiles in registers and applies every weight fragment it
//      fetches to the two of them.
// Estimated shared-memory traffic per tile: ~1.1 MB -> ~0.48 MB.
// =====================================================================================================================
constexpr int PSP = 488;                                                   // pixel pitch (floats) of the P buffer: 61 n-tiles x 8

struct Smem2 {
    static constexpr int C = 64;
    static constexpr int XS = 0;                                       // [XR*XR][C] bf16, swizzled as in v1
    static constexpr int PS = XS + XR * XR * C * 2;                    // [32][PSP] fp32
    static constexpr int SEG = PS + 32 * PSP * 4;
    static constexpr int ACT = SEG + (SR * SR + 1) * 16;
    static constexpr int W2 = ACT + (AR * AR + 1) * 16;
    static constexpr int W3 = W2 + 8 * K3P * 2;
    static constexpr int TOTAL = W3 + 2 * C * K3P * 2;                 // 160 KB: one CTA per SM
};

struct Params2 {
    Params b;             // b.w1 here is the [32][C] "weights stationary" layout: row tap * label_nc + c, rows >= 9 * label_nc zero
    int label_nc;
    int tiles_x, tiles_y, ntiles;
};

__global__ void __launch_bounds__(256, 1) spade_fused_fwd_v2_kernel(const Params2 pp) {
    using S = Smem2;
    constexpr int C = 64, CH = C / 8;
    const Params& p = pp.b;
    extern __shared__ __align__(128) uint8_t smem[];
    float* ps = reinterpret_cast<float*>(smem + S::PS);
    bf16* segs = reinterpret_cast<bf16*>(smem + S::SEG);
    bf16* acts = reinterpret_cast<bf16*>(smem + S::ACT);
    bf16* w2s = reinterpret_cast<bf16*>(smem + S::W2);
    bf16* w3s = reinterpret_cast<bf16*>(smem + S::W3);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int L = pp.label_nc;

    // ---- once per CTA: mlp_shared / gamma|beta operands -> shared memory; x2map operand -> registers (A fragments) ----
    for (int i = tid; i < (8 + 2 * C) * 10; i += 256) {
        const int row = i / 10, k = i - row * 10;
        const bf16* src = row < 8 ? p.w2 + (size_t)row * 80 : p.w3 + (size_t)(row - 8) * 80;
        bf16* dst = row < 8 ? w2s + (size_t)row * K3P : w3s + (size_t)(row - 8) * K3P;
        *reinterpret_cast<uint4*>(dst + k * 8) = __ldg(reinterpret_cast<const uint4*>(src) + k);
    }
    if (tid < 2) *reinterpret_cast<uint4*>(tid == 0 ? segs + SR * SR * 8 : acts + AR * AR * 8) = make_uint4(0u, 0u, 0u, 0u);
    // A fragment of mma.m16n8k16 (row-major 16 x 16): a0 = (row g, k 2t..2t+1), a1 = (row g + 8, same k), a2 = (row g, k + 8), a3 = (row g + 8, k + 8)
    uint32_t wa[2][C / 16][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int kc = 0; kc < C / 16; ++kc) {
            const bf16* r0 = p.w1 + (size_t)(m * 16 + g) * C + kc * 16 + 2 * t4;
            const bf16* r1 = r0 + 8 * C;
            wa[m][kc][0] = __ldg(reinterpret_cast<const uint32_t*>(r0));
            wa[m][kc][1] = __ldg(reinterpret_cast<const uint32_t*>(r1));
            wa[m][kc][2] = __ldg(reinterpret_cast<const uint32_t*>(r0 + 8));
            wa[m][kc][3] = __ldg(reinterpret_cast<const uint32_t*>(r1 + 8));
        }
    const uint32_t xs_a = smem_addr(smem + S::XS), seg_a = smem_addr(segs), act_a = smem_addr(acts);
    const uint32_t w2_a = smem_addr(w2s), w3_a = smem_addr(w3s);
    const int arow = lane & 15, ahalf = lane >> 4;
    const int bn = lane & 7, bhalf = (lane >> 3) & 1;
    const int per_img = pp.tiles_x * pp.tiles_y;

    for (int tile = blockIdx.x; tile < pp.ntiles; tile += gridDim.x) {
        const int img = tile / per_img, trem = tile - img * per_img;
        const int ty0 = (trem / pp.tiles_x) * TILE, tx0 = (trem % pp.tiles_x) * TILE;
        const bf16* ximg = p.x + (long long)img * p.H * p.W * C;
        __syncthreads();                                                 // the previous tile is done with xs / segs / acts
        for (int i = tid; i < XR * XR * CH; i += 256) {
            const int q = i / CH, k = i - q * CH;
            const int iy = ty0 - 3 + q / XR, ix = tx0 - 3 + q % XR;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) v = __ldg(reinterpret_cast<const uint4*>(ximg + ((long long)iy * p.W + ix) * C) + k);
            *reinterpret_cast<uint4*>(smem + S::XS + (size_t)q * C * 2 + ((k ^ (q & 7)) << 4)) = v;
        }
        __syncthreads();

        // ---- stage A1: P[32][484 pixels] = W1 (registers) x pixels: 61 n-tiles of 8 pixels, B = x rows straight from the tile ----
        for (int nt = warp; nt < (XR * XR + 7) / 8; nt += 8) {
            int q = nt * 8 + bn;                                         // the x pixel whose row address this lane supplies
            if (q >= XR * XR) q = XR * XR - 1;                           // ragged last tile: any valid row, results unused
            float acc[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[m][e] = 0.f;
#pragma unroll
            for (int kc = 0; kc < C / 16; ++kc) {
                uint32_t b[2];
                ldsm_x2(b, xs_a + (uint32_t)q * (C * 2) + ((uint32_t)((2 * kc + bhalf) ^ (q & 7)) << 4));
                mma_bf16(acc[0], wa[0][kc], b);
                mma_bf16(acc[1], wa[1][kc], b);
            }
            // D: rows (tap * L + c) g and g + 8 of each 16-row half, columns = pixels nt * 8 + 2 t4, + 1
            const int px = nt * 8 + 2 * t4;
            if (px < PSP) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    *reinterpret_cast<float2*>(ps + (size_t)(m * 16 + g) * PSP + px) = make_float2(acc[m][0], acc[m][1]);
                    *reinterpret_cast<float2*>(ps + (size_t)(m * 16 + g + 8) * PSP + px) = make_float2(acc[m][2], acc[m][3]);
                }
            }
        }
        __syncthreads();
        // ---- stage A2: seg[pixel][c] = b1[c] + sum over the nine taps of the shifted P values (fixed order, fp32) ----
        for (int pq = tid; pq < SR * SR; pq += 256) {
            const int sy = pq / SR, sx = pq - sy * SR;
            const int iy = ty0 - 2 + sy, ix = tx0 - 2 + sx;
            const bool inside = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (inside) {
                for (int c = 0; c < L; ++c) {
                    float a = 0.f;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) a += ps[(size_t)(tap * L + c) * PSP + (sy + tap / 3) * XR + sx + tap % 3];
                    v[c] = a + __ldg(p.b1 + c);
                }
            }
            const uint4 o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), 0u, 0u);
            *reinterpret_cast<uint4*>(segs + pq * 8) = o;
            if (inside && sy >= 2 && sy < 2 + TILE && sx >= 2 && sx < 2 + TILE)
                *reinterpret_cast<uint4*>(p.seg + (((long long)img * p.H + iy) * p.W + ix) * 8) = o;
        }
        __syncthreads();

        // ---- stage B: identical to v1 ----
        for (int mt = warp; mt < (AR * AR + 15) / 16; mt += 8) {
            const int pa = mt * 16 + arow;
            const int aay = pa / AR, aax = pa - aay * AR;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int tap = 2 * j + ahalf;
                const int q = (pa < AR * AR && tap < 9) ? (aay + tap / 3) * SR + aax + tap % 3 : SR * SR;
                uint32_t a[4], b[2];
                ldsm_x4(a, seg_a + (uint32_t)q * 16);
                ldsm_x2(b, w2_a + (uint32_t)(bn * K3P + j * 16 + bhalf * 8) * 2);
                mma_bf16(acc, a, b);
            }
            const float bia0 = __ldg(p.b2 + 2 * t4), bia1 = __ldg(p.b2 + 2 * t4 + 1);
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int pq = mt * 16 + g + 8 * hrow;
                if (pq >= AR * AR) continue;
                const int ay = pq / AR, ax = pq - ay * AR;
                const int iy = ty0 - 1 + ay, ix = tx0 - 1 + ax;
                const bool inside = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                const uint32_t v = inside ? pack_bf16(fmaxf(acc[2 * hrow] + bia0, 0.f), fmaxf(acc[2 * hrow + 1] + bia1, 0.f)) : 0u;
                *reinterpret_cast<uint32_t*>(acts + pq * 8 + 2 * t4) = v;
                if (inside && ay >= 1 && ay < 1 + TILE && ax >= 1 && ax < 1 + TILE)
                    *reinterpret_cast<uint32_t*>(p.actv + (((long long)img * p.H + iy) * p.W + ix) * 8 + 2 * t4) = v;
            }
        }
        __syncthreads();

        // ---- stage C: the warp's two M-tiles (warp, warp + 8) share every weight fragment ----
        uint32_t a2[2][5][4];
        int oy[2][2], ox[2][2];
        bool ok[2][2];
        long long gpix[2][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int mt = warp + 8 * u;
            const int pa = mt * 16 + arow;
            const int oay = pa / TILE, oax = pa - oay * TILE;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int tap = 2 * j + ahalf;
                const int q = tap < 9 ? (oay + tap / 3) * AR + oax + tap % 3 : AR * AR;
                ldsm_x4(a2[u][j], act_a + (uint32_t)q * 16);
            }
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int pq = mt * 16 + g + 8 * hrow;
                oy[u][hrow] = pq / TILE; ox[u][hrow] = pq - oy[u][hrow] * TILE;
                const int iy = ty0 + oy[u][hrow], ix = tx0 + ox[u][hrow];
                ok[u][hrow] = iy < p.H && ix < p.W;
                gpix[u][hrow] = ((long long)img * p.H + iy) * p.W + ix;
            }
        }
#pragma unroll 1
        for (int cj = 0; cj < C / 8; ++cj) {
            float ag[2][4], ab[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int e = 0; e < 4; ++e) { ag[u][e] = 0.f; ab[u][e] = 0.f; }
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                uint32_t bgm[2], bbt[2];
                ldsm_x2(bgm, w3_a + (uint32_t)((cj * 8 + bn) * K3P + j * 16 + bhalf * 8) * 2);
                ldsm_x2(bbt, w3_a + (uint32_t)((C + cj * 8 + bn) * K3P + j * 16 + bhalf * 8) * 2);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    mma_bf16(ag[u], a2[u][j], bgm);
                    mma_bf16(ab[u], a2[u][j], bbt);
                }
            }
            const int ch = cj * 8 + 2 * t4;
            const float g0 = __ldg(p.b3 + ch), g1 = __ldg(p.b3 + ch + 1), e0 = __ldg(p.b3 + C + ch), e1 = __ldg(p.b3 + C + ch + 1);
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    if (!ok[u][hrow]) continue;
                    const float gm0 = bf16_round(ag[u][2 * hrow] + g0), gm1 = bf16_round(ag[u][2 * hrow + 1] + g1);
                    const float bt0 = bf16_round(ab[u][2 * hrow] + e0), bt1 = bf16_round(ab[u][2 * hrow + 1] + e1);
                    const int q = (oy[u][hrow] + 3) * XR + ox[u][hrow] + 3;
                    const uint32_t xv = *reinterpret_cast<const uint32_t*>(smem + S::XS + (size_t)q * C * 2 + ((cj ^ (q & 7)) << 4) + t4 * 4);
                    const float x0 = __uint_as_float(xv << 16), x1 = __uint_as_float(xv & 0xffff0000u);
                    *reinterpret_cast<uint32_t*>(p.y + gpix[u][hrow] * C + ch) = pack_bf16(fmaf(x0, 1.f + gm0, bt0), fmaf(x1, 1.f + gm1, bt1));
                    if (p.gb) {
                        *reinterpret_cast<uint32_t*>(p.gb + gpix[u][hrow] * (2 * C) + ch) = pack_bf16(gm0, gm1);
                        *reinterpret_cast<uint32_t*>(p.gb + gpix[u][hrow] * (2 * C) + C + ch) = pack_bf16(bt0, bt1);
                    }
                }
        }
    }
}

static int launch_v2(const Params2& pp, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        SSG_CHECK_CUDA(cudaFuncSetAttribute(spade_fused_fwd_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem2::TOTAL));
        attr = true;
    }
    int grid = sm_count_cached();
    if (grid > pp.ntiles) grid = pp.ntiles;
    spade_fused_fwd_v2_kernel<<<grid, 256, Smem2::TOTAL, st>>>(pp);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // namespace spf
}  // namespace ssg
using namespace ssg;

extern "C" {

int ssg_spade_fused_supported(int c, int label_nc, int hidden) {
    return ((c == 64 || c == 128) && label_nc >= 1 && label_nc <= 8 && hidden >= 1 && hidden <= 8) ? 1 : 0;
}

int ssg_spade_fused_fwd(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const void* w3, const float* b3,
                        void* seg, void* actv, void* gb, void* y, int n, int h, int w, int c, ssg_stream_t s) {
    SSG_CHECK_ARG(x && w1 && b1 && w2 && b2 && w3 && b3 && seg && actv && y && n > 0 && h > 0 && w > 0 && n <= 65535 && (h + 15) / 16 <= 65535,
                  "spade_fused_fwd: bad arguments");
    spf::Params p;
    p.x = (const bf16*)x; p.w1 = (const bf16*)w1; p.w2 = (const bf16*)w2; p.w3 = (const bf16*)w3; p.b1 = b1; p.b2 = b2; p.b3 = b3;
    p.seg = (bf16*)seg; p.actv = (bf16*)actv; p.gb = (bf16*)gb; p.y = (bf16*)y; p.N = n; p.H = h; p.W = w;
    if (c == 64) return spf::launch<64>(p, (cudaStream_t)s);
    if (c == 128) return spf::launch<128>(p, (cudaStream_t)s);
    set_error("spade_fused_fwd: C = %d is not built (64 or 128)", c);
    return SSG_ERR_UNSUPPORTED;
}

/* v2 (see the block comment above the kernel): c == 64, label_nc <= 3; w1 is the [32][c] weights-stationary layout. */
int ssg_spade_fused_fwd_v2(const void* x, const void* w1m, const float* b1, const void* w2, const float* b2, const void* w3, const float* b3,
                           void* seg, void* actv, void* gb, void* y, int n, int h, int w, int c, int label_nc, ssg_stream_t s) {
    SSG_CHECK_ARG(x && w1m && b1 && w2 && b2 && w3 && b3 && seg && actv && y && n > 0 && h > 0 && w > 0, "spade_fused_fwd_v2: bad arguments");
    if (c != 64 || label_nc < 1 || label_nc > 3) {
        set_error("spade_fused_fwd_v2: built for C = 64, label_nc <= 3 (got %d, %d)", c, label_nc);
        return SSG_ERR_UNSUPPORTED;
    }
    spf::Params2 pp;
    pp.b.x = (const bf16*)x; pp.b.w1 = (const bf16*)w1m; pp.b.w2 = (const bf16*)w2; pp.b.w3 = (const bf16*)w3;
    pp.b.b1 = b1; pp.b.b2 = b2; pp.b.b3 = b3; pp.b.seg = (bf16*)seg; pp.b.actv = (bf16*)actv; pp.b.gb = (bf16*)gb; pp.b.y = (bf16*)y;
    pp.b.N = n; pp.b.H = h; pp.b.W = w; pp.label_nc = label_nc;
    pp.tiles_x = (w + spf::TILE - 1) / spf::TILE; pp.tiles_y = (h + spf::TILE - 1) / spf::TILE;
    const long long nt = (long long)n * pp.tiles_x * pp.tiles_y;
    SSG_CHECK_ARG(nt < (1LL << 31), "spade_fused_fwd_v2: too many tiles");
    pp.ntiles = (int)nt;
    return spf::launch_v2(pp, (cudaStream_t)s);
}

}  // extern "C"
