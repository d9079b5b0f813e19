// Halo-tile tcgen05 implicit-GEMM convolution (stride 1, 1x1 / 3x3 same-size, bf16 NHWC, fp32 accumulate in TMEM).
//
// The plain kernel (conv_tc.cu) re-reads a shifted 128-pixel A tile for every filter tap: 16 KB of A + BN*128 B of
// weights per 64-channel k-block, i.e. >= 96 B/clk/SM at full tensor rate while the L2 delivers ~42 B/clk/SM.  This
// kernel cuts the operand traffic per MMA cycle three ways:
//   * ONE halo tile per (M-tile, 64-channel chunk): an 18 x 10 pixel TMA box (zero-filled outside the image) feeds all
//     nine taps.  Tap (r, s) is the same shared memory seen through a UMMA descriptor whose start address is advanced
//     by (r * 10 + s) pixels and whose stride-byte-offset is the halo pitch (10 * 128 B): the SWIZZLE_128B pattern is a
//     function of absolute smem address bits, so any 128-byte row may start a K-major operand (verified on B200 with
//     scratch/halo_test.cu).  A traffic drops 9 * 128 / 180 = 6.4x.
//   * MT M-tiles (16 x 8 pixels each) per CTA share every weight tile: B traffic per MMA cycle drops MT x.
//   * persistent CTAs (one per SM) walk (M super-tile, N tile) work items with the N tile fastest, so CTAs running
//     together read the same activations from L2; the 2 x MT x BN = 512 TMEM columns hold two accumulator sets, so the
//     epilogue of one work item overlaps the MMAs of the next.
// Warp roles (256 threads): warp 0 = A (halo) TMA producer, warp 1 = B (weights) TMA producer, warp 2 = MMA issuer +
// TMEM owner, warp 3 idle, warps 4..7 = epilogue (TMEM -> registers -> bias / activation -> bf16 -> global).
#include "tc_common.cuh"
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

namespace ssg {
namespace tc {

int encode_bf16_map(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, CUtensorMapSwizzle swizzle, const uint32_t* elem_strides);

constexpr int H_TH = 16, H_TW = 8;              // M-tile: 16 rows x 8 columns = 128 pixels = UMMA M
constexpr int H_THREADS = 256;
constexpr int H_A_TILE_STRIDE = 23552;          // 18 * 10 * 128 B = 23040 rounded up to 1024

struct HaloParams {
    bf16* y;
    const float* bias;
    int bias_n;
    int N, H, W;                 // image dims (output == input: stride 1, same-size)
    int cout;                    // stored output channels (row stride of y)
    int tiles_x, tiles_y;        // M-tiles per image
    int m_tiles;                 // N * tiles_y * tiles_x
    int n_tiles;                 // ceil(cout / BN)
    int total_items;             // supers * n_tiles
    int supers;                  // ceil(m_tiles / MT)
    int n_major;                 // item order: 0 = N tile fastest (CTAs running together share activations in L2);
                                 // 1 = M super-tile fastest (a CTA keeps one N tile's weights resident, RES mode)
    int chunks0, chunks1;        // 64-channel chunks from x0 / x1 (virtual concat)
    int k_last;                  // 16-channel K steps (1..4) that hold live channels in the LAST chunk: thin inputs
                                 // (8..48 stored channels) issue 1..3 MMAs per tap instead of 4 over TMA zero fill
    int ntaps, ksize;
    int flip;                    // 0: tap t reads halo (t / k, t % k) (forward); 1: (k-1 - t / k, k-1 - t % k) (data gradient)
    int halo_c, halo_r;          // halo box dims in pixels: (8 + k - 1) x (16 + k - 1)
    int org;                     // halo origin = tile origin - org
    int act;
    float slope;
    double* stats;               // optional [2][cout] fp64 (pre-zeroed): per-channel sum / sum of squares of the stored y
    int accumulate;              // 1: y += result (TMA reduce-add) instead of y = result
    int split_c;                 // > 0 (multiple of 64): output channels [0, split_c) go to tmY, [split_c, cout) to tmY1 -- the
                                 // data gradient of a virtually concatenated input lands in its two source tensors
    int issuers;                 // 1 or 2 MMA-issuing warps (2: the item's MT accumulators are split between warps 2 and 3)
    int lean;                    // resident-weights instances, 3 x 3, one chunk: the straight-line issue path (halo_issue_res3)
};

// ACCS accumulator sets (2: epilogue overlaps the next item's MMAs).  RES: single-chunk (Cin <= 64) convolutions keep
// all nine weight tiles of the current N tile resident in shared memory instead of streaming them per work item.
template <int MT, int BN, int ACCS, bool RES, int EW = 1>
struct HaloCfg {
    static constexpr int THREADS = 128 + 128 * EW;         // four role warps + EW epilogue warp groups
    static constexpr int OUT_BYTES = EW * (4 * 4096 + 4 * 512);   // epilogue staging: one 32-pixel x 64-channel bf16 box per warp
                                                                  // + one BN-float bias row per warp
    static constexpr int BUDGET = 227 * 1024 - 256 - 1024 - OUT_BYTES;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int A_STAGE_BYTES = MT * H_A_TILE_STRIDE;
    static constexpr int B_MIN = RES ? 9 : 4;
    static constexpr int A_STAGES_RAW = (BUDGET - B_MIN * B_BYTES) / A_STAGE_BYTES;
    static constexpr int A_STAGES = RES ? (A_STAGES_RAW > 6 ? 6 : A_STAGES_RAW) : (A_STAGES_RAW >= 3 ? 3 : 2);
    static constexpr int B_SLOTS_RAW = (BUDGET - A_STAGES * A_STAGE_BYTES) / B_BYTES;
    static constexpr int B_SLOTS = RES ? 9 : (B_SLOTS_RAW > 6 ? 6 : B_SLOTS_RAW);
    static constexpr int B_OFFSET = A_STAGES * A_STAGE_BYTES;
    static constexpr int OUT_OFFSET = B_OFFSET + B_SLOTS * B_BYTES;
    static constexpr int BAR_OFFSET = OUT_OFFSET + OUT_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
    static constexpr int ACC_COLS = MT * BN;           // TMEM columns of one accumulator set
    static_assert(ACCS * ACC_COLS <= 512, "accumulator sets must fit in TMEM");
    static_assert(EW == 1 || (EW == 2 && MT % 2 == 0), "two epilogue groups split an item's M-tiles");
    static_assert(B_SLOTS >= 2 && B_SLOTS <= B_SLOTS_RAW && A_STAGES >= 2, "operand rings do not fit");
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

// MMA-issue loop of conv_tc_halo_kernel for the M-tiles [mt0, mt0 + MTI) of every work item (MTI == MT: the only issuer).
template <int MT, int BN, int ACCS, bool RES, int EW, int MTI>
__device__ __forceinline__ void halo_issue(uint8_t* smem, uint64_t* a_full, uint64_t* a_empty, uint64_t* b_full, uint64_t* b_empty,
                                           uint64_t* acc_full, uint64_t* acc_empty, uint64_t* b_free, uint32_t tmem_base,
                                           const HaloParams& p, int chunks, int mt0) {
    using C = HaloCfg<MT, BN, ACCS, RES, EW>;
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t a_hi = desc_hi((uint32_t)p.halo_c * 128, 2), b_hi = desc_hi(1024, 2);
    const uint32_t b_lo_base = desc_lo(smem_u32(smem + C::B_OFFSET), 16);
    // taps in weight order t = r * k + s; the halo offset advances by one pixel per s and one halo row per r
    // (mirrored for the data gradient) -- no division or table lookup on the issue path
    const int step_s = p.flip ? -8 : 8, step_r = (p.flip ? -1 : 1) * (p.halo_c - p.ksize) * 8;
    const uint32_t a_first = p.flip ? (uint32_t)((p.ksize - 1) * p.halo_c + p.ksize - 1) * 8 : 0u;        // 8 x 16 B per pixel
    int ac = 0, bc = 0, it = 0, cur_n0 = -1, reloads = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
        const int acc = it % ACCS;
        bool fresh_b = !RES;            // RES: the resident weight tiles are awaited once, by the first item that uses them
        if (RES) {
            const int n0 = (p.n_major ? item / p.supers : item % p.n_tiles) * BN;
            if (n0 != cur_n0) { cur_n0 = n0; ++reloads; fresh_b = true; }
        }
        mbar_wait(&acc_empty[acc], ((it / ACCS) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(acc * C::ACC_COLS + mt0 * BN);
        for (int ch = 0; ch < chunks; ++ch, ++ac) {
            const int st = ac % C::A_STAGES;
            const int ksteps = ch == chunks - 1 ? p.k_last : 4;
            mbar_wait(&a_full[st], (ac / C::A_STAGES) & 1);
            tc_fence_after();
            uint32_t a_lo = desc_lo(smem_u32(smem + st * C::A_STAGE_BYTES + mt0 * H_A_TILE_STRIDE), 16) + a_first;
            int tq = 0;
            for (int tap = 0; tap < p.ntaps; ++tap, ++bc) {
                const int sl = RES ? tap : bc % C::B_SLOTS;
                if (fresh_b) {
                    mbar_wait(&b_full[sl], RES ? ((reloads - 1) & 1) : ((bc / C::B_SLOTS) & 1));
                    tc_fence_after();
                }
                const uint32_t b_lo = b_lo_base + (uint32_t)(sl * (C::B_BYTES / 16));
                const uint32_t keep = (uint32_t)(ch | tap);                                         // 0: first k-block of the item
                // k outer, M-tile inner: consecutive MMAs accumulate into DIFFERENT TMEM accumulators
                if (ksteps == 4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
#pragma unroll
                        for (int mt = 0; mt < MTI; ++mt)
                            umma_bf16_lohi_pred(d_base + (uint32_t)(mt * BN), a_lo + (uint32_t)(mt * (H_A_TILE_STRIDE / 16) + 2 * k), a_hi,
                                                b_lo + (uint32_t)(2 * k), b_hi, idesc, k == 0 ? keep : 1u, leader);
                    }
                } else {
                    for (int k = 0; k < ksteps; ++k) {
#pragma unroll
                        for (int mt = 0; mt < MTI; ++mt)
                            umma_bf16_lohi_pred(d_base + (uint32_t)(mt * BN), a_lo + (uint32_t)(mt * (H_A_TILE_STRIDE / 16) + 2 * k), a_hi,
                                                b_lo + (uint32_t)(2 * k), b_hi, idesc, k == 0 ? keep : 1u, leader);
                    }
                }
                if (!RES) umma_commit_pred(&b_empty[sl], leader);
                a_lo += (uint32_t)step_s;
                if (++tq == p.ksize) { tq = 0; a_lo += (uint32_t)step_r; }
            }
            umma_commit_pred(&a_empty[st], leader);
        }
        umma_commit_pred(&acc_full[acc], leader);
        if (RES) {
            // last item on these weights?  then tell the producer when its MMAs are done
            const int nxt = item + (int)gridDim.x;
            const bool last_use = nxt >= p.total_items || (p.n_major ? nxt / p.supers : nxt % p.n_tiles) * BN != cur_n0;
            if (last_use) umma_commit_pred(b_free, leader);
        }
    }
}

// Lean issue path of the streaming (non-resident) instances: the tap loop is unrolled over the KS x KS compile-time halo offsets, the
// ring positions are carried as (slot, phase) pairs instead of being re-derived by division, and full 64-channel chunks take a
// branch-free body.  Same protocol as halo_issue (one b_full wait and one b_empty commit per weight tile).
template <int MT, int BN, int ACCS, int EW, int MTI, int KS>
__device__ __forceinline__ void halo_issue_stream(uint8_t* smem, uint64_t* a_full, uint64_t* a_empty, uint64_t* b_full, uint64_t* b_empty,
                                                  uint64_t* acc_full, uint64_t* acc_empty, uint32_t tmem_base, const HaloParams& p,
                                                  int chunks, int mt0) {
    using C = HaloCfg<MT, BN, ACCS, false, EW>;
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    constexpr int HC = H_TW + KS - 1;
    constexpr int LAST = ((KS - 1) * HC + KS - 1) * 8;
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t a_hi = desc_hi((uint32_t)HC * 128, 2), b_hi = desc_hi(1024, 2);
    const uint32_t b_lo_base = desc_lo(smem_u32(smem + C::B_OFFSET), 16);
    const uint32_t a_lo_base = desc_lo(smem_u32(smem + mt0 * H_A_TILE_STRIDE), 16) + (p.flip ? (uint32_t)LAST : 0u);
    const int sgn = p.flip ? -1 : 1;
    int ast = 0, bsl = 0, acc = 0;
    uint32_t aph = 0, bph = 0, cph = 1;                  // phases of the A ring, the B ring and the accumulator-empty barriers
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        mbar_wait(&acc_empty[acc], cph);
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(acc * C::ACC_COLS + mt0 * BN);
        for (int ch = 0; ch < chunks; ++ch) {
            const bool full = ch != chunks - 1 || p.k_last == 4;
            mbar_wait(&a_full[ast], aph);
            tc_fence_after();
            const uint32_t a_st = a_lo_base + (uint32_t)(ast * (C::A_STAGE_BYTES / 16));
            uint32_t a_row = a_st;
#pragma unroll 1
            for (int r = 0; r < KS; ++r, a_row += (uint32_t)(sgn * HC * 8)) {      // filter rows rolled (code size), columns unrolled
#pragma unroll
                for (int sx = 0; sx < KS; ++sx) {
                    mbar_wait(&b_full[bsl], bph);
                    tc_fence_after();
                    const uint32_t a_lo = a_row + (uint32_t)(sgn * sx * 8);
                    const uint32_t b_lo = b_lo_base + (uint32_t)(bsl * (C::B_BYTES / 16));
                    const uint32_t keep = sx ? 1u : (uint32_t)(ch | r);                            // 0: first k-block of the item
                    if (full) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
#pragma unroll
                            for (int mt = 0; mt < MTI; ++mt)
                                umma_bf16_lohi_pred(d_base + (uint32_t)(mt * BN), a_lo + (uint32_t)(mt * (H_A_TILE_STRIDE / 16) + 2 * k), a_hi,
                                                    b_lo + (uint32_t)(2 * k), b_hi, idesc, k == 0 ? keep : 1u, leader);
                        }
                    } else {
                        for (int k = 0; k < p.k_last; ++k) {
#pragma unroll
                            for (int mt = 0; mt < MTI; ++mt)
                                umma_bf16_lohi_pred(d_base + (uint32_t)(mt * BN), a_lo + (uint32_t)(mt * (H_A_TILE_STRIDE / 16) + 2 * k), a_hi,
                                                    b_lo + (uint32_t)(2 * k), b_hi, idesc, k == 0 ? keep : 1u, leader);
                        }
                    }
                    umma_commit_pred(&b_empty[bsl], leader);
                    if (++bsl == C::B_SLOTS) { bsl = 0; bph ^= 1u; }
                }
            }
            umma_commit_pred(&a_empty[ast], leader);
            if (++ast == C::A_STAGES) { ast = 0; aph ^= 1u; }
        }
        umma_commit_pred(&acc_full[acc], leader);
        if (++acc == ACCS) { acc = 0; cph ^= 1u; }
    }
}

// Lean issue path of the resident-weights instances (one 64-channel chunk, 3 x 3 taps): the nine tap offsets, the K steps of the
// chunk and the weight slots are compile-time, the wait for the resident weights is peeled out of the tap loop, and the body is
// straight-line code -- 9 * KSTEPS * MTI MMAs per item with one add per descriptor.  The generic loop above spends ~280 cycles per tap
// on its branches and barrier bookkeeping (ncu source view, profiles/r02_ncu_halo_thin_in_dconv0_fwd.txt): for Cin = 8 that is
// 2.5 K cycles per item around 0.6 K tensor-clocks of MMAs, for Cin = 64 it equals the MMAs' own 2.3 K tensor-clocks.
template <int MT, int BN, int ACCS, int EW, int MTI, int KSTEPS>
__device__ __forceinline__ void halo_issue_res3(uint8_t* smem, uint64_t* a_full, uint64_t* a_empty, uint64_t* b_full, uint64_t* acc_full,
                                                uint64_t* acc_empty, uint64_t* b_free, uint32_t tmem_base, const HaloParams& p, int mt0) {
    using C = HaloCfg<MT, BN, ACCS, true, EW>;
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    constexpr int HC = H_TW + 2;                                   // halo pitch in pixels
    constexpr int LAST = (2 * HC + 2) * 8;                         // descriptor offset (16-byte units) of halo pixel (2, 2)
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t a_hi = desc_hi((uint32_t)HC * 128, 2), b_hi = desc_hi(1024, 2);
    const uint32_t b_lo_base = desc_lo(smem_u32(smem + C::B_OFFSET), 16);
    const uint32_t a_lo_base = desc_lo(smem_u32(smem + mt0 * H_A_TILE_STRIDE), 16) + (p.flip ? (uint32_t)LAST : 0u);
    const int sgn = p.flip ? -1 : 1;
    int it = 0, cur_n0 = -1, reloads = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
        const int acc = it % ACCS, st = it % C::A_STAGES;
        const int n0 = (p.n_major ? item / p.supers : item % p.n_tiles) * BN;
        if (n0 != cur_n0) {                                        // new N tile: its nine weight tiles have been (re)loaded
            cur_n0 = n0; ++reloads;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) mbar_wait(&b_full[tap], (reloads - 1) & 1);
        }
        mbar_wait(&acc_empty[acc], ((it / ACCS) & 1) ^ 1);
        mbar_wait(&a_full[st], (it / C::A_STAGES) & 1);
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(acc * C::ACC_COLS + mt0 * BN);
        const uint32_t a_st = a_lo_base + (uint32_t)(st * (C::A_STAGE_BYTES / 16));
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a_lo = a_st + (uint32_t)(sgn * (((tap / 3) * HC + tap % 3) * 8));
            const uint32_t b_lo = b_lo_base + (uint32_t)(tap * (C::B_BYTES / 16));
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
#pragma unroll
                for (int mt = 0; mt < MTI; ++mt)
                    umma_bf16_lohi_pred(d_base + (uint32_t)(mt * BN), a_lo + (uint32_t)(mt * (H_A_TILE_STRIDE / 16) + 2 * k), a_hi,
                                        b_lo + (uint32_t)(2 * k), b_hi, idesc, (tap | k) ? 1u : 0u, leader);
            }
        }
        umma_commit_pred(&a_empty[st], leader);
        umma_commit_pred(&acc_full[acc], leader);
        const int nxt = item + (int)gridDim.x;                     // last item on these weights?  then tell the producer when its MMAs are done
        if (nxt >= p.total_items || (p.n_major ? nxt / p.supers : nxt % p.n_tiles) * BN != cur_n0) umma_commit_pred(b_free, leader);
    }
}

template <int MT, int BN, int ACCS, int EW, int MTI>
__device__ __forceinline__ void halo_issue_res3_k(uint8_t* smem, uint64_t* a_full, uint64_t* a_empty, uint64_t* b_full, uint64_t* acc_full,
                                                  uint64_t* acc_empty, uint64_t* b_free, uint32_t tmem_base, const HaloParams& p, int mt0) {
    switch (p.k_last) {
        case 1: halo_issue_res3<MT, BN, ACCS, EW, MTI, 1>(smem, a_full, a_empty, b_full, acc_full, acc_empty, b_free, tmem_base, p, mt0); break;
        case 2: halo_issue_res3<MT, BN, ACCS, EW, MTI, 2>(smem, a_full, a_empty, b_full, acc_full, acc_empty, b_free, tmem_base, p, mt0); break;
        case 3: halo_issue_res3<MT, BN, ACCS, EW, MTI, 3>(smem, a_full, a_empty, b_full, acc_full, acc_empty, b_free, tmem_base, p, mt0); break;
        default: halo_issue_res3<MT, BN, ACCS, EW, MTI, 4>(smem, a_full, a_empty, b_full, acc_full, acc_empty, b_free, tmem_base, p, mt0); break;
    }
}

template <int MT, int BN, int ACCS, bool RES, int EW>
__global__ void __launch_bounds__(128 + 128 * EW, 1) conv_tc_halo_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                                          const __grid_constant__ CUtensorMap tmA1,
                                                                          const __grid_constant__ CUtensorMap tmB,
                                                                          const __grid_constant__ CUtensorMap tmY,
                                                                          const __grid_constant__ CUtensorMap tmY1, const HaloParams p) {
    using C = HaloCfg<MT, BN, ACCS, RES, EW>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFFSET);
    uint64_t* a_empty = a_full + C::A_STAGES;
    uint64_t* b_full = a_empty + C::A_STAGES;
    uint64_t* b_empty = b_full + C::B_SLOTS;
    uint64_t* acc_full = b_empty + C::B_SLOTS;
    uint64_t* acc_empty = acc_full + ACCS;
    uint64_t* b_free = acc_empty + ACCS;         // RES: the MMAs that read the resident weights have completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_free + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = p.chunks0 + p.chunks1;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (threadIdx.x == 0) {
        const int ni = (p.issuers == 2 && MT >= 2) ? 2 : 1;     // MMA-issuing warps: arrivals per consumer-side barrier
        for (int i = 0; i < C::A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], ni); }
        for (int i = 0; i < C::B_SLOTS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], ni); }
        for (int i = 0; i < ACCS; ++i) { mbar_init(&acc_full[i], ni); mbar_init(&acc_empty[i], 128 * EW); }
        mbar_init(b_free, ni);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);     // warp-uniform (uniform datapath in the MMA warp)

    if (warp == 0) {
        // ---- A producer: one halo tile per (M-tile, chunk) ----
        if (lane == 0) {
            tma_prefetch_desc(&tmA0);
            const uint32_t a_bytes = (uint32_t)(MT * p.halo_c * p.halo_r * 128);
            int ac = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int super = p.n_major ? item % p.supers : item / p.n_tiles;
                for (int ch = 0; ch < chunks; ++ch, ++ac) {
                    const int st = ac % C::A_STAGES;
                    mbar_wait(&a_empty[st], ((ac / C::A_STAGES) & 1) ^ 1);
                    mbar_expect_tx(&a_full[st], a_bytes);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const int t = super * MT + mt;          // tiles past m_tiles have img >= N: the box is zero-filled
                        const int img = t / tiles_per_img, rem = t - img * tiles_per_img;
                        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
                        uint8_t* dst = smem + st * C::A_STAGE_BYTES + mt * H_A_TILE_STRIDE;
                        if (ch < p.chunks0) tma_load_4d(dst, &tmA0, ch * 64, tx * H_TW - p.org, ty * H_TH - p.org, img, &a_full[st]);
                        else tma_load_4d(dst, &tmA1, (ch - p.chunks0) * 64, tx * H_TW - p.org, ty * H_TH - p.org, img, &a_full[st]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---- B producer: one weight tile per (chunk, tap) ----
        if (lane == 0) {
            tma_prefetch_desc(&tmB);
            int bc = 0, cur_n0 = -1, reloads = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int n0 = (p.n_major ? item / p.supers : item % p.n_tiles) * BN;
                if (RES) {
                    if (n0 == cur_n0) continue;                  // weights of this N tile are resident
                    if (reloads > 0) mbar_wait(b_free, (reloads - 1) & 1);
                    cur_n0 = n0; ++reloads;
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        mbar_expect_tx(&b_full[tap], C::B_BYTES);
                        tma_load_3d(smem + C::B_OFFSET + tap * C::B_BYTES, &tmB, 0, n0, tap, &b_full[tap]);
                    }
                    continue;
                }
                for (int ch = 0; ch < chunks; ++ch)
                    for (int tap = 0; tap < p.ntaps; ++tap, ++bc) {
                        const int sl = bc % C::B_SLOTS;
                        mbar_wait(&b_empty[sl], ((bc / C::B_SLOTS) & 1) ^ 1);
                        mbar_expect_tx(&b_full[sl], C::B_BYTES);
                        tma_load_3d(smem + C::B_OFFSET + sl * C::B_BYTES, &tmB, ch * 64, n0, tap, &b_full[sl]);
                    }
            }
        }
    } else if (warp == 2 || (warp == 3 && p.issuers == 2 && MT >= 2)) {
        // ---- MMA issue: the WHOLE warp walks the (warp-uniform) loop so that barrier addresses, descriptors and TMEM
        // addresses live on the uniform datapath; the elected lane alone issues tcgen05.mma / tcgen05.commit.
        // With p.issuers == 2 the MT accumulators of an item are split between warps 2 and 3: MMAs into different accumulators
        // are independent, so the two issue streams need no ordering; each commits its own MMAs to the rings' empty barriers
        // (initialised with one arrival per issuer).  The thin instances (BN = 16: 8 tensor-clocks per MMA) and the BN = 64
        // instances (32 clocks) were paced by ONE warp's descriptor arithmetic on the uniform datapath, not by the tensor pipe.
        bool done = false;
        if constexpr (RES) {
            if (p.lean) {
                if (p.issuers == 2 && MT >= 2)
                    halo_issue_res3_k<MT, BN, ACCS, EW, (MT >= 2 ? MT / 2 : 1)>(smem, a_full, a_empty, b_full, acc_full, acc_empty, b_free, tmem_base,
                                                                                 p, warp == 2 ? 0 : MT / 2);
                else
                    halo_issue_res3_k<MT, BN, ACCS, EW, MT>(smem, a_full, a_empty, b_full, acc_full, acc_empty, b_free, tmem_base, p, 0);
                done = true;
            }
        }
        if constexpr (!RES) {
            if (p.lean) {
                constexpr int MTI2 = MT >= 2 ? MT / 2 : 1;
                const bool two = p.issuers == 2 && MT >= 2;
                const int m0 = (two && warp == 3) ? MT / 2 : 0;
                if (p.ksize == 3) {
                    if (two) halo_issue_stream<MT, BN, ACCS, EW, MTI2, 3>(smem, a_full, a_empty, b_full, b_empty, acc_full, acc_empty, tmem_base, p, chunks, m0);
                    else halo_issue_stream<MT, BN, ACCS, EW, MT, 3>(smem, a_full, a_empty, b_full, b_empty, acc_full, acc_empty, tmem_base, p, chunks, 0);
                } else {
                    if (two) halo_issue_stream<MT, BN, ACCS, EW, MTI2, 1>(smem, a_full, a_empty, b_full, b_empty, acc_full, acc_empty, tmem_base, p, chunks, m0);
                    else halo_issue_stream<MT, BN, ACCS, EW, MT, 1>(smem, a_full, a_empty, b_full, b_empty, acc_full, acc_empty, tmem_base, p, chunks, 0);
                }
                done = true;
            }
        }
        if (done) {
        } else if (p.issuers == 2 && MT >= 2) {
            halo_issue<MT, BN, ACCS, RES, EW, (MT >= 2 ? MT / 2 : 1)>(smem, a_full, a_empty, b_full, b_empty, acc_full, acc_empty, b_free, tmem_base, p,
                                                                   chunks, warp == 2 ? 0 : MT / 2);
        } else {
            halo_issue<MT, BN, ACCS, RES, EW, MT>(smem, a_full, a_empty, b_full, b_empty, acc_full, acc_empty, b_free, tmem_base, p, chunks, 0);
        }
    } else if (warp >= 4) {
        // ---- epilogue: warp q owns TMEM lanes [32q, 32q + 32) = tile rows 4q .. 4q + 3 (a {64 ch, 8, 4, 1} box of y).
        // TMEM -> registers -> bias / activation -> bf16 -> swizzled smem staging -> ONE TMA store per (tile, 64 channels):
        // full-line coalesced writes, clipped to the image / channel bounds by the TMA unit.
        // EW == 2: a second group of four warps (8..11, the same TMEM lane quarters) takes the odd M-tiles of every item -- with
        // Cin <= 64 an item's MMAs (<= 2.3 K tensor-clocks) are shorter than ONE warp group's drain of its 2 x 128 x 64 outputs
        // (~4 K clocks of TMEM load -> convert -> stage -> TMA store latency), so the epilogue paced those layers.
        const int q = warp & 3, grp = EW == 2 ? ((warp - 4) >> 2) : 0, slot = grp * 4 + q;
        uint8_t* stage = smem + C::OUT_OFFSET + slot * 4096;
        float* bias_s = reinterpret_cast<float*>(smem + C::OUT_OFFSET + EW * 4 * 4096 + slot * 512);
        uint8_t* my_row = stage + lane * 128;
        const int sw = lane & 7;
        const bool has_bias = p.bias != nullptr;
        // act(v) = max(v, v * neg): neg = 1 (identity), 0 (ReLU) or the LeakyReLU slope -- branch-free in the hot loop
        const float neg = p.act == SSG_ACT_RELU ? 0.f : (p.act == SSG_ACT_LEAKY ? p.slope : 1.f);
        // BatchNorm statistics of the output (batchnorm.py:59-64) as a by-product: lane l keeps fp32 partial sums of
        // channels 2l, 2l+1 of each 64-channel half over every row this warp stores; flushed with fp64 atomics when the
        // N tile changes.  The values summed are the bf16-rounded ones the normalisation will read back.
        const bool want_stats = p.stats != nullptr;
        constexpr int NH = (BN + 63) / 64;                     // 64-channel halves of the N tile
        float st[NH][4];
#pragma unroll
        for (int i = 0; i < NH; ++i) st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
        int stats_n0 = -1;
        auto flush_stats = [&]() {
#pragma unroll
            for (int i = 0; i < NH; ++i) {
                const int c = stats_n0 + i * 64 + 2 * lane;
                if (c < p.cout) { atomicAdd(p.stats + c, (double)st[i][0]); atomicAdd(p.stats + p.cout + c, (double)st[i][2]); }
                if (c + 1 < p.cout) { atomicAdd(p.stats + c + 1, (double)st[i][1]); atomicAdd(p.stats + p.cout + c + 1, (double)st[i][3]); }
                st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
            }
        };
        int it = 0, bias_n0 = -1;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
            const int acc = it % ACCS;
            const int super = p.n_major ? item % p.supers : item / p.n_tiles;
            const int n0 = (p.n_major ? item / p.supers : item % p.n_tiles) * BN;
            if (has_bias && n0 != bias_n0) {                        // stage this N tile's bias row (zeros past bias_n)
                __syncwarp();
#pragma unroll
                for (int e = lane; e < BN; e += 32) bias_s[e] = (n0 + e < p.bias_n) ? __ldg(p.bias + n0 + e) : 0.f;
                __syncwarp();
                bias_n0 = n0;
            }
            if (want_stats && n0 != stats_n0) {
                if (stats_n0 >= 0) flush_stats();
                stats_n0 = n0;
            }
            mbar_wait(&acc_full[acc], (it / ACCS) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int mt = grp; mt < MT; mt += EW) {
                const int t = super * MT + mt;
                if (t >= p.m_tiles) break;
                const int img = t / tiles_per_img, rem = t - img * tiles_per_img;
                const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C::ACC_COLS + mt * BN);
#pragma unroll
                for (int h0 = 0; h0 < BN; h0 += 64) {
                    if (n0 + h0 >= p.cout) break;
                    // thin outputs (BN = 16: SPADE's x2map, the logits head, data gradients into 8-channel maps): only the 16 live
                    // accumulator columns are read, converted and staged -- the store's other channels are clipped by the TMA
                    // unit anyway (a full 64-column step per M-tile made these instances epilogue-bound: 4 steps per item)
                    constexpr int CW = BN < 64 ? BN : 64, NJ = CW / 8;
                    uint32_t v[64];
                    if constexpr (CW == 64) {
                        tmem_ld_32x32b_x32(t_addr + (uint32_t)h0, v);
                        tmem_ld_32x32b_x32(t_addr + (uint32_t)(h0 + 32), v + 32);
                    } else {
                        static_assert(CW == 16, "narrow epilogue: BN = 16");
                        tmem_ld_32x32b_x16(t_addr + (uint32_t)h0, v);
                    }
                    tmem_ld_wait();
                    if (has_bias) {
#pragma unroll
                        for (int e = 0; e < CW; e += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + h0 + e);
                            v[e] = __float_as_uint(__uint_as_float(v[e]) + b4.x);
                            v[e + 1] = __float_as_uint(__uint_as_float(v[e + 1]) + b4.y);
                            v[e + 2] = __float_as_uint(__uint_as_float(v[e + 2]) + b4.z);
                            v[e + 3] = __float_as_uint(__uint_as_float(v[e + 3]) + b4.w);
                        }
                    }
                    if (lane == 0) tma_store_wait_read();          // the previous store has finished reading the staging box
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {                  // 8 channels = one 16-byte chunk
                        uint32_t w4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float a = __uint_as_float(v[j * 8 + 2 * e]), b = __uint_as_float(v[j * 8 + 2 * e + 1]);
                            a = fmaxf(a, a * neg);
                            b = fmaxf(b, b * neg);
                            const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
                            w4[e] = *reinterpret_cast<const uint32_t*>(&hh);
                        }
                        *reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        const int cg = n0 + h0;
                        const bool second = p.split_c > 0 && cg >= p.split_c;
                        const CUtensorMap* ymap = second ? &tmY1 : &tmY;
                        const int cy = second ? cg - p.split_c : cg;
                        if (p.accumulate) tma_reduce_add_4d(ymap, stage, cy, tx * H_TW, ty * H_TH + 4 * q, img);
                        else tma_store_4d(ymap, stage, cy, tx * H_TW, ty * H_TH + 4 * q, img);
                        tma_store_commit();
                    }
                    if (want_stats) {
                        const uint32_t valid = __ballot_sync(0xffffffffu, (ty * H_TH + 4 * q + (lane >> 3) < p.H) && (tx * H_TW + (lane & 7) < p.W));
                        float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
                        const uint8_t* col = stage + ((lane & 3) << 2);
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            if ((valid >> r) & 1u) {                 // warp-uniform: rows outside the image are skipped
                                const uint32_t w2 = *reinterpret_cast<const uint32_t*>(col + r * 128 + (((lane >> 2) ^ (r & 7)) << 4));
                                const float a = __uint_as_float(w2 << 16), b = __uint_as_float(w2 & 0xffff0000u);
                                a0 += a; a1 += b;
                                q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
                            }
                        }
                        st[h0 / 64][0] += a0; st[h0 / 64][1] += a1; st[h0 / 64][2] += q0; st[h0 / 64][3] += q1;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
        }
        if (want_stats && stats_n0 >= 0) flush_stats();
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int MT, int BN, int ACCS, bool RES, int EW = 1>
static int launch_halo(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& ym, const CUtensorMap& ym1,
                       HaloParams& p, cudaStream_t st) {
    using C = HaloCfg<MT, BN, ACCS, RES, EW>;
    static bool attr_set = false;
    if (!attr_set) {
        SSG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_halo_kernel<MT, BN, ACCS, RES, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
        attr_set = true;
    }
    p.n_tiles = (p.cout + BN - 1) / BN;
    p.supers = (p.m_tiles + MT - 1) / MT;
    p.total_items = p.supers * p.n_tiles;
    p.n_major = RES ? 1 : 0;
    // Two issuing warps wherever an item has >= 2 accumulators (measured: thin BN = 16 instances +20 %, BN = 64 +5-10 %, BN = 128
    // +6 %); SSG_HALO_ISSUERS=1 restores the single issuer (A/B switch)
    static const int issuers_env = getenv("SSG_HALO_ISSUERS") ? atoi(getenv("SSG_HALO_ISSUERS")) : 2;
    p.issuers = (MT >= 2 && issuers_env != 1) ? 2 : 1;
    static const bool lean_off = getenv("SSG_HALO_LEAN") && atoi(getenv("SSG_HALO_LEAN")) == 0;       // A/B switch
    static const bool lean_stream_off = getenv("SSG_HALO_LEAN_STREAM") && atoi(getenv("SSG_HALO_LEAN_STREAM")) == 0;
    if (RES) p.lean = (p.ksize == 3 && p.chunks0 + p.chunks1 == 1 && !lean_off) ? 1 : 0;
    else p.lean = ((p.ksize == 3 || p.ksize == 1) && !lean_off && !lean_stream_off) ? 1 : 0;
    int grid = sm_count_cached();
    if (grid > p.total_items) grid = p.total_items;
    conv_tc_halo_kernel<MT, BN, ACCS, RES, EW><<<grid, C::THREADS, C::TOTAL, st>>>(a0, a1, b, ym, ym1, p);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// Same-size stride-1 convolution (flip = 0) / data gradient (flip = 1: tap t reads the mirrored halo offset).
int run_conv_halo(const void* x0, int c0, const void* x1, int c1, const void* w_packed, int w_taps, const float* bias, int bias_n,
                  void* y, int n, int h, int w, int gemm_n, int ksize, int flip, int act, float slope, double* stats, cudaStream_t st,
                  int accumulate, void* y1, int split_c) {
    HaloParams p;
    memset(&p, 0, sizeof(p));
    p.stats = stats;
    p.accumulate = accumulate;
    p.split_c = y1 ? split_c : 0;
    SSG_CHECK_ARG(!y1 || (split_c > 0 && split_c % 64 == 0 && split_c < gemm_n && (gemm_n - split_c) % 8 == 0 && !stats),
                  "conv halo: split output needs split_c %% 64 == 0 and 0 < split_c < cout (split_c=%d cout=%d)", split_c, gemm_n);
    if (stats) SSG_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)gemm_n, st));
    p.y = (bf16*)y; p.bias = bias; p.bias_n = bias_n; p.N = n; p.H = h; p.W = w; p.cout = gemm_n;
    p.tiles_x = (w + H_TW - 1) / H_TW; p.tiles_y = (h + H_TH - 1) / H_TH; p.m_tiles = n * p.tiles_x * p.tiles_y;
    p.chunks0 = (c0 + 63) / 64; p.chunks1 = (c1 + 63) / 64;
    {
        const int last = c1 > 0 ? c1 - 64 * (p.chunks1 - 1) : c0 - 64 * (p.chunks0 - 1);     // live channels of the last chunk
        p.k_last = (last + 15) / 16;
    }
    p.ntaps = ksize * ksize; p.ksize = ksize; p.flip = flip;
    p.halo_c = H_TW + ksize - 1; p.halo_r = H_TH + ksize - 1; p.org = (ksize - 1) / 2;
    p.act = act; p.slope = slope;
    CUtensorMap ma0, ma1, mb;
    auto enc_act = [&](CUtensorMap* m, const void* ptr, int c) {
        uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        uint64_t str[3] = {(uint64_t)c * 2, (uint64_t)w * c * 2, (uint64_t)h * w * c * 2};
        uint32_t box[4] = {64, (uint32_t)p.halo_c, (uint32_t)p.halo_r, 1};
        return encode_bf16_map(m, ptr, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, nullptr);
    };
    int rc = enc_act(&ma0, x0, c0);
    if (rc) return rc;
    ma1 = ma0;
    if (c1 > 0) {
        rc = enc_act(&ma1, x1, c1);
        if (rc) return rc;
    }
    const int cin = c0 + c1;
    // tile shape: SSG_HALO_SHAPE = "464" (4 M-tiles x BN 64) or "2128" (2 M-tiles x BN 128), two accumulator sets each; default by Cout
    static const char* shape_env = getenv("SSG_HALO_SHAPE");
    int shape = gemm_n > 64 ? 2128 : (gemm_n <= 16 ? 416 : 464);      // thin outputs (SPADE maps, logits): N = 16 MMAs
    if (shape_env && gemm_n > 64) shape = atoi(shape_env);
    // single-chunk inputs (Cin <= 64: level-0 layers, stems, SPADE's thin maps): all nine weight tiles stay resident
    static const bool no_res = getenv("SSG_HALO_NORES") != nullptr;
    if (c0 + c1 <= 64 && gemm_n > 16 && !no_res) shape = 264;
    const bool wide = shape == 2128;
    const int bn = wide ? 128 : (shape == 416 ? 16 : 64);
    {
        uint64_t dims[3] = {(uint64_t)cin, (uint64_t)gemm_n, (uint64_t)w_taps};
        uint64_t str[2] = {(uint64_t)cin * 2, (uint64_t)gemm_n * cin * 2};
        uint32_t box[3] = {64, (uint32_t)bn, 1};
        rc = encode_bf16_map(&mb, w_packed, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, nullptr);
        if (rc) return rc;
    }
    CUtensorMap my, my1;
    auto enc_out = [&](CUtensorMap* m, void* ptr, int c) {
        uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        uint64_t str[3] = {(uint64_t)c * 2, (uint64_t)w * c * 2, (uint64_t)h * w * c * 2};
        uint32_t box[4] = {64, H_TW, 4, 1};
        return encode_bf16_map(m, ptr, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, nullptr);
    };
    rc = enc_out(&my, y, p.split_c ? p.split_c : gemm_n);
    if (rc) return rc;
    my1 = my;
    if (p.split_c) {
        rc = enc_out(&my1, y1, gemm_n - p.split_c);
        if (rc) return rc;
    }
    static const char* res_env = getenv("SSG_HALO_RES");          // "264" (default) or "164"
    if (shape == 264 && res_env && atoi(res_env) == 164) return launch_halo<1, 64, 4, true>(ma0, ma1, mb, my, my1, p, st);
    // Epilogue warp groups of the resident-weights instance.  Thin inputs (<= 32 channels: SPADE's gamma|beta convolution, the
    // discriminator's first layer) issue <= 18 short MMAs per item and are paced by the drain of the 2 x 128 x 64 outputs: two
    // groups (0.240 -> 0.162 ms on 16 x 512^2 x 8 -> 64).  With 64 input channels the second group's instructions compete with the
    // MMA-issuing warps for the same schedulers and the kernel gets SLOWER (0.274 -> 0.310 ms): one group.
    // (Two groups on the 128-wide streaming instance cost it an operand stage and lose: 0.130 -> 0.172 ms on a 1 x 1 768 -> 256 at 128^2.)
    // SSG_HALO_EPI = 1: always one group, 22: always two (A/B switches)
    static const int epi_env = getenv("SSG_HALO_EPI") ? atoi(getenv("SSG_HALO_EPI")) : 2;
    if (shape == 264 && epi_env == 2 && p.k_last <= 2) return launch_halo<2, 64, 2, true, 2>(ma0, ma1, mb, my, my1, p, st);
    if (shape == 264 && epi_env == 22) return launch_halo<2, 64, 2, true, 2>(ma0, ma1, mb, my, my1, p, st);
    if (shape == 264) return launch_halo<2, 64, 2, true>(ma0, ma1, mb, my, my1, p, st);
    if (shape == 2128) return launch_halo<2, 128, 2, false>(ma0, ma1, mb, my, my1, p, st);
    if (shape == 416) return launch_halo<4, 16, 2, false>(ma0, ma1, mb, my, my1, p, st);
    return launch_halo<4, 64, 2, false>(ma0, ma1, mb, my, my1, p, st);
}

// =====================================================================================================
// Data gradient of the 3 x 3 / stride-2 / pad-1 convolutions (the discriminator's down-sampling layers, models_seg_gan.py:38-39)
// on the halo formulation.  dx[2i + py, 2j + px] = sum over the taps (r, s) with r == 1 (py == 0) or r in {0, 2} (py == 1), same
// for s / px, of dy[i + (r == 0), j + (s == 0)] W[r][s]: every tap belongs to exactly ONE output-parity class, and all four
// classes read the same 17 x 9 pixel halo box of dy.  A work item = (16 x 8 tile of dy positions, 64-wide tile of dx channels):
// ONE TMA box of dy per 64-channel chunk feeds nine MMA groups into FOUR accumulators (one per parity class, 4 x 64 TMEM columns;
// two sets), and the epilogue stores each class through its own tensor map -- dx seen as {C, W/2, H/2, N} with doubled pixel
// strides and the base advanced by (py, px) -- so every store is full 128-byte lines.  The plain kernel (conv_tc.cu) loads a shifted
// dy tile per tap and class (9 x 16 KB + 9 weight tiles per 128 positions: L2-bandwidth bound, profiles/r02_ncu_s2_dgrad_merged_l0.txt).
// Warps: 0 = dy producer, 1 = weight producer, 2 = issuer of class (1,1)'s four taps, 3 = issuer of the other five taps,
// 4..11 = two epilogue groups (classes {0, 2} and {1, 3}).  Optional mask: the producer layer's activation backward (act'(out)).
// =====================================================================================================
struct S2DgradParams {
    int N, OH, OW;               // dy dims (tile space)
    int H, W;                    // dx dims (2 * OH, 2 * OW)
    int cin;                     // dx channels: GEMM N extent
    int tiles_x, tiles_y, m_tiles, n_tiles, total_items;
    int chunks, k_last;          // 64-channel chunks of dy (GEMM K) and live 16-channel K steps of the last one
    int res;                     // chunks == 1: the nine weight tiles of an N tile stay resident, items walk M fastest
    const bf16* mask;            // optional, shaped like dx
    int mask_act;
    float mask_slope;
};

struct S2DgradCfg {
    static constexpr int THREADS = 384;
    static constexpr int A_STAGE_BYTES = 20480;             // 17 x 9 x 128 B = 19584, rounded up to 1024
    static constexpr int A_BOX_BYTES = 17 * 9 * 128;
    static constexpr int A_STAGES = 5;
    static constexpr int B_BYTES = 64 * 128;
    static constexpr int B_OFFSET = A_STAGES * A_STAGE_BYTES;
    static constexpr int OUT_OFFSET = B_OFFSET + 9 * B_BYTES;
    static constexpr int OUT_BYTES = 8 * 4096;
    static constexpr int BAR_OFFSET = OUT_OFFSET + OUT_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

// tap t = 3 r + s: parity class (r != 1) * 2 + (s != 1), halo offset ((r == 0) * 9 + (s == 0)) pixels, first tap of its class?
__device__ __forceinline__ constexpr int s2d_class(int t) { return ((t / 3) != 1 ? 2 : 0) + ((t % 3) != 1 ? 1 : 0); }
__device__ __forceinline__ constexpr int s2d_aoff(int t) { return (((t / 3) == 0 ? 9 : 0) + ((t % 3) == 0 ? 1 : 0)) * 8; }
__device__ __forceinline__ constexpr bool s2d_first(int t) { return t == 0 || t == 1 || t == 3 || t == 4; }

template <int WHICH>    // 0: the four taps of class (1,1); 1: the other five
__device__ __forceinline__ void s2d_issue(uint8_t* smem, uint64_t* a_full, uint64_t* a_empty, uint64_t* b_full, uint64_t* b_empty,
                                          uint64_t* acc_full, uint64_t* acc_empty, uint64_t* b_free, uint32_t tmem_base,
                                          const S2DgradParams& p) {
    using C = S2DgradCfg;
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t a_hi = desc_hi(9 * 128, 2), b_hi = desc_hi(1024, 2);
    const uint32_t b_lo_base = desc_lo(smem_u32(smem + C::B_OFFSET), 16);
    const uint32_t a_lo_base = desc_lo(smem_u32(smem), 16);
    const int supers = p.m_tiles;
    int ast = 0, acc = 0, cur_n0 = -1, reloads = 0;
    uint32_t aph = 0, cph = 1, bph = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const int n0 = (p.res ? item / supers : item % p.n_tiles) * 64;
        bool fresh = !p.res;
        if (p.res && n0 != cur_n0) { cur_n0 = n0; ++reloads; fresh = true; }
        mbar_wait(&acc_empty[acc], cph);
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(acc * 256);
        for (int ch = 0; ch < p.chunks; ++ch) {
            const bool full = ch != p.chunks - 1 || p.k_last == 4;
            mbar_wait(&a_full[ast], aph);
            tc_fence_after();
            const uint32_t a_st = a_lo_base + (uint32_t)(ast * (C::A_STAGE_BYTES / 16));
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                if ((s2d_class(t) == 3) != (WHICH == 0)) continue;
                if (fresh) {
                    mbar_wait(&b_full[t], p.res ? (uint32_t)((reloads - 1) & 1) : bph);
                    tc_fence_after();
                }
                const uint32_t a_lo = a_st + (uint32_t)s2d_aoff(t);
                const uint32_t b_lo = b_lo_base + (uint32_t)(t * (C::B_BYTES / 16));
                const uint32_t d = d_base + (uint32_t)(s2d_class(t) * 64);
                const uint32_t keep = s2d_first(t) ? (uint32_t)ch : 1u;
                if (full) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_lohi_pred(d, a_lo + (uint32_t)(2 * k), a_hi, b_lo + (uint32_t)(2 * k), b_hi, idesc, k == 0 ? keep : 1u, leader);
                } else {
                    for (int k = 0; k < p.k_last; ++k)
                        umma_bf16_lohi_pred(d, a_lo + (uint32_t)(2 * k), a_hi, b_lo + (uint32_t)(2 * k), b_hi, idesc, k == 0 ? keep : 1u, leader);
                }
                if (!p.res) umma_commit_pred(&b_empty[t], leader);
            }
            umma_commit_pred(&a_empty[ast], leader);
            if (++ast == C::A_STAGES) { ast = 0; aph ^= 1u; }
            bph ^= 1u;
        }
        umma_commit_pred(&acc_full[acc], leader);
        if (++acc == 2) { acc = 0; cph ^= 1u; }
        if (p.res) {
            const int nxt = item + (int)gridDim.x;
            if (nxt >= p.total_items || (nxt / supers) * 64 != cur_n0) umma_commit_pred(b_free, leader);
        }
    }
}

__global__ void __launch_bounds__(S2DgradCfg::THREADS, 1) conv_tc_halo_s2dgrad_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                                        const __grid_constant__ CUtensorMap tmB,
                                                                                        const __grid_constant__ CUtensorMap tmY0,
                                                                                        const __grid_constant__ CUtensorMap tmY1,
                                                                                        const __grid_constant__ CUtensorMap tmY2,
                                                                                        const __grid_constant__ CUtensorMap tmY3,
                                                                                        const S2DgradParams p) {
    using C = S2DgradCfg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFFSET);
    uint64_t* a_empty = a_full + C::A_STAGES;
    uint64_t* b_full = a_empty + C::A_STAGES;
    uint64_t* b_empty = b_full + 9;
    uint64_t* acc_full = b_empty + 9;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* b_free = acc_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_free + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int supers = p.m_tiles;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 2); }
        for (int i = 0; i < 9; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 2); mbar_init(&acc_empty[i], 256); }
        mbar_init(b_free, 2);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tmA);
            int ast = 0;
            uint32_t aph = 1;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int t = p.res ? item % supers : item / p.n_tiles;
                const int img = t / tiles_per_img, rem = t - img * tiles_per_img;
                const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
                for (int ch = 0; ch < p.chunks; ++ch) {
                    mbar_wait(&a_empty[ast], aph);
                    mbar_expect_tx(&a_full[ast], C::A_BOX_BYTES);
                    tma_load_4d(smem + ast * C::A_STAGE_BYTES, &tmA, ch * 64, tx * H_TW, ty * H_TH, img, &a_full[ast]);
                    if (++ast == C::A_STAGES) { ast = 0; aph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            tma_prefetch_desc(&tmB);
            int cur_n0 = -1, reloads = 0;
            uint32_t bph = 1;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int n0 = (p.res ? item / supers : item % p.n_tiles) * 64;
                if (p.res) {
                    if (n0 == cur_n0) continue;
                    if (reloads > 0) mbar_wait(b_free, (uint32_t)((reloads - 1) & 1));
                    cur_n0 = n0; ++reloads;
                    for (int t = 0; t < 9; ++t) {
                        mbar_expect_tx(&b_full[t], C::B_BYTES);
                        tma_load_3d(smem + C::B_OFFSET + t * C::B_BYTES, &tmB, 0, n0, t, &b_full[t]);
                    }
                    continue;
                }
                for (int ch = 0; ch < p.chunks; ++ch, bph ^= 1u)
                    for (int t = 0; t < 9; ++t) {
                        mbar_wait(&b_empty[t], bph);
                        mbar_expect_tx(&b_full[t], C::B_BYTES);
                        tma_load_3d(smem + C::B_OFFSET + t * C::B_BYTES, &tmB, ch * 64, n0, t, &b_full[t]);
                    }
            }
        }
    } else if (warp == 2) {
        s2d_issue<0>(smem, a_full, a_empty, b_full, b_empty, acc_full, acc_empty, b_free, tmem_base, p);
    } else if (warp == 3) {
        s2d_issue<1>(smem, a_full, a_empty, b_full, b_empty, acc_full, acc_empty, b_free, tmem_base, p);
    } else {
        // ---- epilogue: group g = classes {g, g + 2} (px = g, py = 0 / 1); warp q of a group owns tile rows 4q .. 4q + 3 ----
        const int q = warp & 3, grp = (warp - 4) >> 2, slot = grp * 4 + q;
        uint8_t* stage = smem + C::OUT_OFFSET + slot * 4096;
        uint8_t* my_row = stage + lane * 128;
        const int sw = lane & 7;
        int acc = 0;
        uint32_t fph = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            const int t = p.res ? item % supers : item / p.n_tiles;
            const int n0 = (p.res ? item / supers : item % p.n_tiles) * 64;
            const int img = t / tiles_per_img, rem = t - img * tiles_per_img;
            const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
            const int oy = ty * H_TH + 4 * q + (lane >> 3), ox = tx * H_TW + (lane & 7);       // this lane's dy position
            const bool pix_ok = oy < p.OH && ox < p.OW;
            // the producer's output rows of BOTH classes are requested before the wait for the accumulator: their DRAM latency
            // (the epilogue's only global loads) then overlaps the item's MMAs instead of stalling every store step
            uint4 mk[2][8];
            const bool use_mask = p.mask != nullptr && pix_ok;
            if (use_mask) {
#pragma unroll
                for (int py = 0; py < 2; ++py) {
                    const bf16* mrow = p.mask + ((((long long)img * p.H + (2 * oy + py)) * p.W + (2 * ox + grp)) * p.cin + n0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        mk[py][j] = (n0 + j * 8 < p.cin) ? __ldg(reinterpret_cast<const uint4*>(mrow + j * 8)) : make_uint4(0, 0, 0, 0);
                }
            }
            mbar_wait(&acc_full[acc], fph);
            tc_fence_after();
#pragma unroll
            for (int py = 0; py < 2; ++py) {
                const int cls = py * 2 + grp;
                uint32_t v[64];
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + cls * 64);
                tmem_ld_32x32b_x32(t_addr, v);
                tmem_ld_32x32b_x32(t_addr + 32u, v + 32);
                tmem_ld_wait();
                if (use_mask) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (n0 + j * 8 < p.cin) {
                            const uint4 m4 = mk[py][j];
                            const uint32_t mw[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float m0 = __uint_as_float(mw[e] << 16), m1 = __uint_as_float(mw[e] & 0xffff0000u);
                                v[j * 8 + 2 * e] = __float_as_uint(__uint_as_float(v[j * 8 + 2 * e]) * act_grad_from_out(m0, p.mask_act, p.mask_slope));
                                v[j * 8 + 2 * e + 1] =
                                    __float_as_uint(__uint_as_float(v[j * 8 + 2 * e + 1]) * act_grad_from_out(m1, p.mask_act, p.mask_slope));
                            }
                        }
                    }
                }
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t w4[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(v[j * 8 + 2 * e]), __uint_as_float(v[j * 8 + 2 * e + 1]));
                        w4[e] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    *reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    const CUtensorMap* ymap = cls == 0 ? &tmY0 : (cls == 1 ? &tmY1 : (cls == 2 ? &tmY2 : &tmY3));
                    tma_store_4d(ymap, stage, n0, tx * H_TW, ty * H_TH + 4 * q, img);
                    tma_store_commit();
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
            if (++acc == 2) { acc = 0; fph ^= 1u; }
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

bool dgrad_s2_halo_supported(int h, int w, int cin, int cout) {
    static const bool off = getenv("SSG_S2_HALO") && atoi(getenv("SSG_S2_HALO")) == 0;        // A/B switch: the plain kernel
    return !off && h % 2 == 0 && w % 2 == 0 && h >= 2 && w >= 2 && cin % 8 == 0 && cout % 8 == 0;
}

int run_dgrad_s2_halo(const void* dy, int cout, const void* w_packed, void* dx, int n, int h, int w, int cin, const void* mask,
                      int mask_act, float mask_slope, cudaStream_t st) {
    using C = S2DgradCfg;
    S2DgradParams p;
    memset(&p, 0, sizeof(p));
    p.N = n; p.H = h; p.W = w; p.OH = h / 2; p.OW = w / 2; p.cin = cin;
    p.tiles_x = (p.OW + H_TW - 1) / H_TW; p.tiles_y = (p.OH + H_TH - 1) / H_TH; p.m_tiles = n * p.tiles_x * p.tiles_y;
    p.n_tiles = (cin + 63) / 64;
    p.total_items = p.m_tiles * p.n_tiles;
    p.chunks = (cout + 63) / 64;
    p.k_last = (cout - 64 * (p.chunks - 1) + 15) / 16;
    p.res = p.chunks == 1 ? 1 : 0;
    p.mask = (const bf16*)mask; p.mask_act = mask_act; p.mask_slope = mask_slope;
    CUtensorMap ma, mb, my[4];
    {
        uint64_t dims[4] = {(uint64_t)cout, (uint64_t)p.OW, (uint64_t)p.OH, (uint64_t)n};
        uint64_t str[3] = {(uint64_t)cout * 2, (uint64_t)p.OW * cout * 2, (uint64_t)p.OH * p.OW * cout * 2};
        uint32_t box[4] = {64, 9, 17, 1};
        int rc = encode_bf16_map(&ma, dy, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, nullptr);
        if (rc) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)cout, (uint64_t)cin, 9};
        uint64_t str[2] = {(uint64_t)cout * 2, (uint64_t)cin * cout * 2};
        uint32_t box[3] = {64, 64, 1};
        int rc = encode_bf16_map(&mb, w_packed, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, nullptr);
        if (rc) return rc;
    }
    for (int cls = 0; cls < 4; ++cls) {
        const int py = cls >> 1, px = cls & 1;
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)p.OW, (uint64_t)p.OH, (uint64_t)n};
        uint64_t str[3] = {(uint64_t)cin * 4, (uint64_t)w * cin * 4, (uint64_t)h * w * cin * 2};
        uint32_t box[4] = {64, H_TW, 4, 1};
        int rc = encode_bf16_map(&my[cls], (const uint8_t*)dx + ((size_t)py * w + px) * cin * 2, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                 nullptr);
        if (rc) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        SSG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_halo_s2dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
        attr_set = true;
    }
    int grid = sm_count_cached();
    if (grid > p.total_items) grid = p.total_items;
    conv_tc_halo_s2dgrad_kernel<<<grid, C::THREADS, C::TOTAL, st>>>(ma, mb, my[0], my[1], my[2], my[3], p);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // namespace tc
}  // namespace ssg

// =====================================================================================================
// Halo-tile weight gradient (3x3, stride 1, same-size): dW[co][ci][r][s] = sum_px dy[px][co] * x[px + (r,s) - pad][ci].
// GEMM view: D[M = 128 = two taps x 64 ci][N = 64 co] += A^T B with the pixel index as K, both operands MN-major
// straight from the NHWC TMA boxes.  A CTA owns ONE 64-channel chunk of x, ONE 64-wide co tile and all nine taps
// (five tap pairs -> five 128 x 64 fp32 accumulators = 320 TMEM columns) and walks a strided subset of the 16 x 8
// pixel tiles (split-K).  Per pixel tile it loads ONE 18 x 10 halo box of x (23 KB) and one dy box (16 KB): the nine
// shifted views of x are UMMA descriptors into the same halo tile (start address advanced by (r * 10 + s) pixels,
// stride-byte-offset = halo pitch), the second tap of a pair is reached through the leading-byte-offset.  The plain
// kernel loads 9 x 16 KB of x per 64-channel chunk instead (4.4 x more operand bytes per MMA cycle).
// =====================================================================================================
namespace ssg {
namespace tc {

struct HaloWgradParams {
    float* dw;                   // OIHW fp32 [cout][cin][3][3], pre-zeroed
    int N, H, W;
    int cout, cin;               // real channel extents of dw
    int tiles_x, tiles_y, m_tiles;
    int chunks0, chunks1;
    int issuers;                 // 1 or 2 MMA-issuing warps (see the kernel)
    int splits0, splits1;        // split-K factors over the pixel tiles of pass 0 / pass 1 (gridDim.z = splits0 + splits1; one pass:
                                 // splits1 == 0).  Pass 0 carries three tap pairs, pass 1 two: 3 : 2 splits equalise the CTAs' run times
};

// BN = 64: one pass, five tap-pair accumulators of 64 columns (320 TMEM columns).  Each 128 x 64 x 16 MMA reads 4 KB of A and
// 2 KB of B from shared memory for 32 tensor-clocks = 192 B/clk against the SM's 128 B/clk port: at most 67 % of the tensor
// peak (ncu, profiles/r02_ncu_halo_wgrad_l0.txt: 44 % tensor-active, the issuing warp waiting on the MMA queue).
// BN = 128 (Cout >= 128): a 128 x 128 x 16 MMA reads 4 + 4 KB for 64 tensor-clocks = 125 B/clk -- inside the port.  Five
// 128-column accumulators would need 640 of the 512 TMEM columns, so the nine taps are covered by TWO passes (blockIdx.z & 1):
// pass 0 = tap pairs 0-2 (384 columns), pass 1 = pairs 3-4 (256 columns); each pass streams its pixel tiles' x halo box and a
// 128-channel dy box (two 64-channel TMA boxes, the second reached through the descriptor's leading-byte offset) once.
template <int BN>
struct HaloWgradCfg {
    static constexpr int XS = 4, DS = (BN == 128) ? 3 : 4;    // ring slots
    static constexpr int DY_BYTES = BN * 256;                // 128 pixels x BN channels x 2 B
    static constexpr int X_OFFSET = 0;
    static constexpr int DY_OFFSET = XS * H_A_TILE_STRIDE;
    static constexpr int BAR_OFFSET = DY_OFFSET + DS * DY_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
    static constexpr int PASSES = (BN == 128) ? 2 : 1;
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

// Issue loop for tap pairs [G_LO, G_HI) whose accumulators start at TMEM column (g - G_BASE) * BN: the whole warp walks the loop
// (uniform datapath), the elected lane issues the MMAs and commits.
template <int BN, int G_LO, int G_HI, int G_BASE>
__device__ __forceinline__ void halo_wgrad_issue(uint8_t* smem, uint64_t* x_full, uint64_t* dy_full, uint64_t* x_empty,
                                                 uint64_t* dy_empty, uint64_t* acc_full, uint32_t tmem_base, int n_iter) {
    using C = HaloWgradCfg<BN>;
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);     // both operands MN-major
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t x_hi = desc_hi(1280, 2), dy_hi = desc_hi(1024, 2);
    const uint32_t x_lo_base = desc_lo(smem_u32(smem + C::X_OFFSET), 0);
    const uint32_t dy_lo_base = desc_lo(smem_u32(smem + C::DY_OFFSET), 16384);    // LBO: the second 64-channel box of a 128-wide co tile
    for (int it = 0; it < n_iter; ++it) {
        const int xs = it % C::XS, ds = it % C::DS;
        mbar_wait(&x_full[xs], (it / C::XS) & 1);
        mbar_wait(&dy_full[ds], (it / C::DS) & 1);
        tc_fence_after();
        const uint32_t xk = x_lo_base + (uint32_t)(xs * (H_A_TILE_STRIDE / 16));
        const uint32_t dk = dy_lo_base + (uint32_t)(ds * (C::DY_BYTES / 16));
        const uint32_t keep = (uint32_t)it;
        // Straight-line issue of the tile's MMAs (k outer, tap pair inner: consecutive MMAs accumulate into different
        // TMEM accumulators); every descriptor is (slot base + compile-time constant), the leader flag predicates the
        // instruction itself.  Pair g = taps 2g and 2g+1 (tap 8 is paired with a dummy second half whose rows are never
        // stored).  16 pixels = two tile rows per MMA: x advances 2 halo rows (160 x 16 B), dy 2 box rows (128 x 16 B).
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int g = G_LO; g < G_HI; ++g) {
                const int ta = 2 * g, tb = g < 4 ? 2 * g + 1 : 8;
                const int off_a = ((ta / 3) * 10 + ta % 3) * 128, off_b = ((tb / 3) * 10 + tb % 3) * 128;
                const uint32_t lbo = g < 4 ? (uint32_t)(off_b - off_a) : 128u;
                // start-address field += off_a / 16; LBO field (bits 16..29) = lbo / 16: both compile-time constants
                umma_bf16_lohi_pred(tmem_base + (uint32_t)((g - G_BASE) * BN), xk + (uint32_t)(k * 160 + ((off_a >> 4) | ((lbo >> 4) << 16))), x_hi,
                                    dk + (uint32_t)(k * 128), dy_hi, idesc, k == 0 ? keep : 1u, leader);
            }
        }
        umma_commit_pred(&x_empty[xs], leader);
        umma_commit_pred(&dy_empty[ds], leader);
    }
    umma_commit_pred(acc_full, leader);
}

template <int BN>
__global__ void __launch_bounds__(H_THREADS, 1) conv_tc_halo_wgrad_kernel(const __grid_constant__ CUtensorMap tmX0,
                                                                           const __grid_constant__ CUtensorMap tmX1,
                                                                           const __grid_constant__ CUtensorMap tmDY,
                                                                           const HaloWgradParams p) {
    using C = HaloWgradCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* x_full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFFSET);
    uint64_t* x_empty = x_full + C::XS;
    uint64_t* dy_full = x_empty + C::XS;
    uint64_t* dy_empty = dy_full + C::DS;
    uint64_t* acc_full = dy_empty + C::DS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.x, co0 = blockIdx.y * BN;
    const int pass = (C::PASSES == 2 && (int)blockIdx.z >= p.splits0) ? 1 : 0;          // BN = 128: tap pairs 0-2 / 3-4
    const int zsplit = pass ? (int)blockIdx.z - p.splits0 : (int)blockIdx.z;
    const int nsplit = pass ? p.splits1 : p.splits0;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int n_iter = (p.m_tiles - zsplit + nsplit - 1) / nsplit;
    // this CTA's tap pairs and its MMA-issuing warps' shares of them
    const int g_lo = (C::PASSES == 2 && pass == 1) ? 3 : 0;
    const int g_hi = (C::PASSES == 2 && pass == 0) ? 3 : 5;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::XS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], p.issuers); }
        for (int i = 0; i < C::DS; ++i) { mbar_init(&dy_full[i], 1); mbar_init(&dy_empty[i], p.issuers); }
        mbar_init(acc_full, p.issuers);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < n_iter; ++it) {
                const int t = zsplit + it * nsplit;
                const int img = t / tiles_per_img, rem = t - img * tiles_per_img;
                const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
                const int sl = it % C::XS;
                mbar_wait(&x_empty[sl], ((it / C::XS) & 1) ^ 1);
                mbar_expect_tx(&x_full[sl], 18 * 10 * 128);
                uint8_t* dst = smem + C::X_OFFSET + sl * H_A_TILE_STRIDE;
                if (chunk < p.chunks0) tma_load_4d(dst, &tmX0, chunk * 64, tx * H_TW - 1, ty * H_TH - 1, img, &x_full[sl]);
                else tma_load_4d(dst, &tmX1, (chunk - p.chunks0) * 64, tx * H_TW - 1, ty * H_TH - 1, img, &x_full[sl]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int it = 0; it < n_iter; ++it) {
                const int t = zsplit + it * nsplit;
                const int img = t / tiles_per_img, rem = t - img * tiles_per_img;
                const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
                const int sl = it % C::DS;
                mbar_wait(&dy_empty[sl], ((it / C::DS) & 1) ^ 1);
                mbar_expect_tx(&dy_full[sl], C::DY_BYTES);
#pragma unroll
                for (int hb = 0; hb < BN / 64; ++hb)
                    tma_load_4d(smem + C::DY_OFFSET + sl * C::DY_BYTES + hb * 16384, &tmDY, co0 + hb * 64, tx * H_TW, ty * H_TH, img, &dy_full[sl]);
            }
        }
    } else if (warp == 2 || (warp == 3 && p.issuers == 2)) {
        // MMA issue.  ncu (profiles/r02_ncu_halo_wgrad_l0.txt): a single issuing warp spent a third of its time on the uniform
        // datapath's dependent descriptor arithmetic (short scoreboard) between UTCHMMAs.  With p.issuers == 2 the CTA's tap-pair
        // accumulators are split between warps 2 and 3: MMAs into different accumulators are independent, so two issue streams
        // need no ordering; each stream commits its own MMAs to the rings' empty barriers (one arrival per issuer).
        if (C::PASSES == 1) {
            if (p.issuers == 2) {
                if (warp == 2) halo_wgrad_issue<BN, 0, 3, 0>(smem, x_full, dy_full, x_empty, dy_empty, acc_full, tmem_base, n_iter);
                else halo_wgrad_issue<BN, 3, 5, 0>(smem, x_full, dy_full, x_empty, dy_empty, acc_full, tmem_base, n_iter);
            } else {
                halo_wgrad_issue<BN, 0, 5, 0>(smem, x_full, dy_full, x_empty, dy_empty, acc_full, tmem_base, n_iter);
            }
        } else if (pass == 0) {
            if (p.issuers == 2) {
                if (warp == 2) halo_wgrad_issue<BN, 0, 2, 0>(smem, x_full, dy_full, x_empty, dy_empty, acc_full, tmem_base, n_iter);
                else halo_wgrad_issue<BN, 2, 3, 0>(smem, x_full, dy_full, x_empty, dy_empty, acc_full, tmem_base, n_iter);
            } else {
                halo_wgrad_issue<BN, 0, 3, 0>(smem, x_full, dy_full, x_empty, dy_empty, acc_full, tmem_base, n_iter);
            }
        } else {
            if (p.issuers == 2) {
                if (warp == 2) halo_wgrad_issue<BN, 3, 4, 3>(smem, x_full, dy_full, x_empty, dy_empty, acc_full, tmem_base, n_iter);
                else halo_wgrad_issue<BN, 4, 5, 3>(smem, x_full, dy_full, x_empty, dy_empty, acc_full, tmem_base, n_iter);
            } else {
                halo_wgrad_issue<BN, 3, 5, 3>(smem, x_full, dy_full, x_empty, dy_empty, acc_full, tmem_base, n_iter);
            }
        }
    } else if (warp >= 4) {
        // ---- flush: fp32 reductions straight from the accumulator registers into dW (OIHW).  (A staged variant -- 32 output
        // channels at a time through the idle operand rings as [co][ci][tap], then contiguous reductions -- was built, verified
        // and measured in round 2: no faster at level 0 and 5-30 % slower on the deep layers, where its two block barriers and
        // the shared-memory round trip outweigh the coalescing.  What the flush costs is its COUNT: see the split-K choice in
        // launch_wgrad_halo.)
        const int q = warp & 3;
        const int row = q * 32 + lane;                  // accumulator row: tap half (row >> 6), ci (row & 63)
        const int ci = chunk * 64 + (row & 63);
        mbar_wait(acc_full, 0);
        tc_fence_after();
        if (n_iter > 0) {
#pragma unroll 1
            for (int g = g_lo; g < g_hi; ++g) {
                const int tap = 2 * g + (row >> 6);
                const bool row_ok = tap < 9 && ci < p.cin;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((g - g_lo) * BN + c0), v);
                    tmem_ld_wait();
                    if (!row_ok) continue;
                    float* dst = p.dw + ((long long)(co0 + c0) * p.cin + ci) * 9 + tap;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (co0 + c0 + j < p.cout) atomicAdd(dst + (long long)j * p.cin * 9, __uint_as_float(v[j]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int BN>
static int launch_wgrad_halo(const CUtensorMap& mx0, const CUtensorMap& mx1, const CUtensorMap& mdy, HaloWgradParams& p, int chunks,
                             int cout_real, cudaStream_t st) {
    using C = HaloWgradCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        SSG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_halo_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
        attr_set = true;
    }
    const int co_tiles = (cout_real + BN - 1) / BN;
    const int pairs = chunks * co_tiles;
    // Split-K factor z = CTAs per (chunk, co tile).  One CTA per SM is resident (shared memory), so CTAs beyond the SM count run
    // as a second wave -- and every CTA ends with the same fixed-size flush of its accumulators (40 960 fp32 reductions into L2,
    // ~22 % of a level-0 launch in the ncu source page: profiles/r02_ncu_halo_wgrad_l0.txt).  Pick the z that minimises
    //     waves(z) * (tiles per CTA + flush), flush expressed in tile times,
    // instead of round 1's fixed "two waves".
    const int sms = sm_count_cached();
    const long long flush_tiles = (BN == 128) ? 40 : 24;
    int best_z = 1;
    long long best_cost = -1;
    const int z_max = (2 * sms + pairs - 1) / pairs + 1;
    for (int zc = C::PASSES; zc <= z_max && zc <= p.m_tiles * C::PASSES; ++zc) {
        const long long waves = ((long long)pairs * zc + sms - 1) / sms;
        const long long per_cta = ((long long)p.m_tiles * C::PASSES + zc - 1) / zc;      // two-pass: z CTAs share 2 x m_tiles tile-passes
        const long long cost = waves * (per_cta + flush_tiles);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_z = zc; }
    }
    static const int z_env = getenv("SSG_WGRAD_WAVES") ? atoi(getenv("SSG_WGRAD_WAVES")) : 0;      // A/B: 2 = round 1's two waves
    int z = z_env == 2 ? (2 * sms + pairs - 1) / pairs : best_z;
    if (C::PASSES == 1) {
        if (z > p.m_tiles) z = p.m_tiles;
        if (z < 1) z = 1;
        p.splits0 = z; p.splits1 = 0;
    } else {
        if (z < 2) z = 2;
        int s0 = (3 * z + 2) / 5;                                 // pass 0: three tap pairs, pass 1: two
        if (s0 < 1) s0 = 1;
        int s1 = z - s0;
        if (s1 < 1) s1 = 1;
        if (s0 > p.m_tiles) s0 = p.m_tiles;
        if (s1 > p.m_tiles) s1 = p.m_tiles;
        p.splits0 = s0; p.splits1 = s1;
    }
    dim3 grid((unsigned)chunks, (unsigned)co_tiles, (unsigned)(p.splits0 + p.splits1));
    conv_tc_halo_wgrad_kernel<BN><<<grid, H_THREADS, C::TOTAL, st>>>(mx0, mx1, mdy, p);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// x0 | x1: stored channel counts c0 / c1 (multiples of 8); dy: stored channels cout_s.  dw must be zeroed by the caller.
int run_wgrad_halo(const void* x0, int c0, const void* x1, int c1, const void* dy, int cout_s, float* dw, int cout_real, int cin_real,
                   int n, int h, int w, cudaStream_t st) {
    HaloWgradParams p;
    memset(&p, 0, sizeof(p));
    p.dw = dw; p.N = n; p.H = h; p.W = w; p.cout = cout_real; p.cin = cin_real;
    p.tiles_x = (w + H_TW - 1) / H_TW; p.tiles_y = (h + H_TH - 1) / H_TH; p.m_tiles = n * p.tiles_x * p.tiles_y;
    p.chunks0 = (c0 + 63) / 64; p.chunks1 = (c1 + 63) / 64;
    static const int issuers_env = getenv("SSG_WGRAD_ISSUERS") ? atoi(getenv("SSG_WGRAD_ISSUERS")) : 2;
    p.issuers = issuers_env == 1 ? 1 : 2;
    CUtensorMap mx0, mx1, mdy;
    auto enc = [&](CUtensorMap* m, const void* ptr, int c, int bw, int bh) {
        uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        uint64_t str[3] = {(uint64_t)c * 2, (uint64_t)w * c * 2, (uint64_t)h * w * c * 2};
        uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
        return encode_bf16_map(m, ptr, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, nullptr);
    };
    int rc = enc(&mx0, x0, c0, 10, 18);
    if (rc) return rc;
    mx1 = mx0;
    if (c1 > 0) {
        rc = enc(&mx1, x1, c1, 10, 18);
        if (rc) return rc;
    }
    rc = enc(&mdy, dy, cout_s, H_TW, H_TH);
    if (rc) return rc;
    const int chunks = p.chunks0 + p.chunks1;
    // SSG_WGRAD_BN=64 forces the one-pass 64-wide kernel everywhere (A/B switch)
    static const bool force64 = getenv("SSG_WGRAD_BN") != nullptr && atoi(getenv("SSG_WGRAD_BN")) == 64;
    if (cout_real >= 128 && cout_s >= 128 && !force64) return launch_wgrad_halo<128>(mx0, mx1, mdy, p, chunks, cout_real, st);
    return launch_wgrad_halo<64>(mx0, mx1, mdy, p, chunks, cout_real, st);
}

// =====================================================================================================
// Thin-input weight gradient (3x3, stride 1, same size, x stored with 8 channels: image stems, SPADE's label / hidden
// maps).  The halo kernel above spends an M = 128 accumulator on two taps x 64 input channels, 56 of which are zero
// fill here -- five MMAs per 16 pixels for 8 live channels.  This kernel puts ALL nine taps into one accumulator:
//   D[m = tap * 8 + ci][n = co] += A^T B,   K = pixels,
// with the A operand assembled as an im2col image in shared memory by nine small TMA boxes {8 ch, 8 w, 16 h} (one per
// tap, shifted by the tap offset, zero-filled outside the image): 128 pixels x 16 B = 2048 B per tap.  Eight consecutive
// pixels x 16 B are exactly one no-swizzle MN-major core matrix (8 K-rows of 8 contiguous M elements), so the UMMA
// descriptor is: no swizzle, LBO (K direction) = 128 B, SBO (M direction) = 2048 B (semantics verified on B200 by
// scratch/nosw_test.cu).  Groups 9..15 of M read whatever follows in shared memory; their accumulator rows are never
// stored.  B = dy box(es) {64 ch, 8, 16} SWIZZLE_128B MN-major as in the halo kernel (two boxes for a 128-wide co tile,
// reached through the LBO).  One MMA per 16 pixels instead of five; x traffic 18 KB instead of a 23 KB zero-filled halo.
// =====================================================================================================
constexpr int TW_STAGES = 4;
constexpr int TW_A_BYTES = 9 * 2048;
constexpr int TW_A_REGION = ((TW_STAGES * TW_A_BYTES + 7 * 2048 + 1023) / 1024) * 1024;   // + readable tail for M groups 9..15

struct ThinWgradParams {
    float* dw;                   // OIHW fp32 [cout][cin][3][3], accumulated into
    int N, H, W;
    int cout, cin;               // real channel extents of dw
    int tiles_x, tiles_y, m_tiles;
};

template <int BN>
__global__ void __launch_bounds__(H_THREADS, 1) conv_tc_thin_wgrad_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                           const __grid_constant__ CUtensorMap tmDY,
                                                                           const ThinWgradParams p) {
    constexpr int DY_BYTES = BN * 256;                 // 128 pixels x BN channels x 2 B
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_a = smem;
    uint8_t* s_dy = smem + TW_A_REGION;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_dy + TW_STAGES * DY_BYTES);
    uint64_t* empty = full + TW_STAGES;
    uint64_t* acc_full = empty + TW_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int co0 = blockIdx.x * BN;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int n_iter = (p.m_tiles - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TW_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tmX);
            tma_prefetch_desc(&tmDY);
            for (int it = 0; it < n_iter; ++it) {
                const int t = blockIdx.y + it * gridDim.y;
                const int img = t / tiles_per_img, rem = t - img * tiles_per_img;
                const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
                const int sl = it % TW_STAGES;
                mbar_wait(&empty[sl], ((it / TW_STAGES) & 1) ^ 1);
                mbar_expect_tx(&full[sl], TW_A_BYTES + DY_BYTES);
                uint8_t* a_dst = s_a + sl * TW_A_BYTES;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap)
                    tma_load_4d(a_dst + tap * 2048, &tmX, 0, tx * H_TW + tap % 3 - 1, ty * H_TH + tap / 3 - 1, img, &full[sl]);
#pragma unroll
                for (int hb = 0; hb < BN / 64; ++hb)
                    tma_load_4d(s_dy + sl * DY_BYTES + hb * 16384, &tmDY, co0 + hb * 64, tx * H_TW, ty * H_TH, img, &full[sl]);
            }
        }
    } else if (warp == 2) {
        constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);     // both operands MN-major
        const uint32_t leader = elect_one() ? 1u : 0u;
        const uint32_t a_hi = desc_hi(2048, 0), dy_hi = desc_hi(1024, 2);
        const uint32_t a_lo_base = desc_lo(smem_u32(s_a), 128);
        const uint32_t dy_lo_base = desc_lo(smem_u32(s_dy), 16384);    // LBO: second 64-channel box of a 128-wide co tile
        for (int it = 0; it < n_iter; ++it) {
            const int sl = it % TW_STAGES;
            mbar_wait(&full[sl], (it / TW_STAGES) & 1);
            tc_fence_after();
            const uint32_t ak = a_lo_base + (uint32_t)(sl * (TW_A_BYTES / 16));
            const uint32_t dk = dy_lo_base + (uint32_t)(sl * (DY_BYTES / 16));
            const uint32_t keep = (uint32_t)it;
#pragma unroll
            for (int k = 0; k < 8; ++k)            // 16 pixels per MMA: A advances 2 core matrices (256 B), dy 16 rows (2048 B)
                umma_bf16_lohi_pred(tmem_base, ak + (uint32_t)(k * 16), a_hi, dk + (uint32_t)(k * 128), dy_hi, idesc, k == 0 ? keep : 1u, leader);
            umma_commit_pred(&empty[sl], leader);
        }
        umma_commit_pred(acc_full, leader);
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int row = q * 32 + lane;                  // accumulator row = tap * 8 + ci
        const int tap = row >> 3, ci = row & 7;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        if (n_iter > 0 && q < 3) {                      // rows 96..127 hold nothing
            const bool row_ok = tap < 9 && ci < p.cin;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                tmem_ld_wait();
                if (!row_ok) continue;
                float* dst = p.dw + ((long long)(co0 + c0) * p.cin + ci) * 9 + tap;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (co0 + c0 + j < p.cout) atomicAdd(dst + (long long)j * p.cin * 9, __uint_as_float(v[j]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
}

template <int BN>
static int launch_thin_wgrad(const CUtensorMap& mx, const CUtensorMap& mdy, const ThinWgradParams& p, int co_tiles, cudaStream_t st) {
    constexpr int TOTAL = TW_A_REGION + TW_STAGES * BN * 256 + 256 + 1024;
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
    static bool attr_set = false;
    if (!attr_set) {
        SSG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_thin_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOTAL));
        attr_set = true;
    }
    // one wave of CTAs (one CTA per SM is resident): a second wave would pay the accumulator flush and the pipeline fill again
    int splits = sm_count_cached() / co_tiles;
    if (splits > p.m_tiles) splits = p.m_tiles;
    if (splits < 1) splits = 1;
    dim3 grid((unsigned)co_tiles, (unsigned)splits);
    conv_tc_thin_wgrad_kernel<BN><<<grid, H_THREADS, TOTAL, st>>>(mx, mdy, p);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// x: stored with exactly 8 channels; dy: stored channels cout_s.  dw is accumulated into (zeroed by the caller when needed).
int run_wgrad_thin(const void* x, const void* dy, int cout_s, float* dw, int cout_real, int cin_real, int n, int h, int w,
                   cudaStream_t st) {
    ThinWgradParams p;
    memset(&p, 0, sizeof(p));
    p.dw = dw; p.N = n; p.H = h; p.W = w; p.cout = cout_real; p.cin = cin_real;
    p.tiles_x = (w + H_TW - 1) / H_TW; p.tiles_y = (h + H_TH - 1) / H_TH; p.m_tiles = n * p.tiles_x * p.tiles_y;
    CUtensorMap mx, mdy;
    {
        uint64_t dims[4] = {8, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        uint64_t str[3] = {16, (uint64_t)w * 16, (uint64_t)h * w * 16};
        uint32_t box[4] = {8, H_TW, H_TH, 1};
        int rc = encode_bf16_map(&mx, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE, nullptr);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)cout_s, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        uint64_t str[3] = {(uint64_t)cout_s * 2, (uint64_t)w * cout_s * 2, (uint64_t)h * w * cout_s * 2};
        uint32_t box[4] = {64, H_TW, H_TH, 1};
        int rc = encode_bf16_map(&mdy, dy, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, nullptr);
        if (rc) return rc;
    }
    if (cout_real > 64) return launch_thin_wgrad<128>(mx, mdy, p, (cout_real + 127) / 128, st);
    return launch_thin_wgrad<64>(mx, mdy, p, (cout_real + 63) / 64, st);
}

}  // namespace tc
}  // namespace ssg
