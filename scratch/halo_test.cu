// Experiment: can a K-major SWIZZLE_128B UMMA descriptor start at a non-1024-aligned row of a TMA-written halo tile,
// with a stride-byte-offset that is not a multiple of 1024 (halo pitch 10 pixels = 1280 B)?
// A = halo tile [18 rows][10 cols][64 ch] loaded by ONE TMA box; B = 64x64 identity; D[m][n] should equal
// halo[(m / 8) + r][(m % 8) + s][n] for every tap (r, s).  Two descriptor variants are tried: base_offset = 0
// and base_offset = (start >> 7) & 7.
#include "../ssunet-gan_b200/csrc/tc_common.cuh"
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <vector>
using namespace ssg::tc;
typedef __nv_bfloat16 bf16;

constexpr int HR = 18, HC = 10, C = 64;

__global__ void __launch_bounds__(128) halo_kernel(const __grid_constant__ CUtensorMap tmA, float* out /*[2][9][128][64]*/) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // 180 * 128 = 23040 B  (-> 23552 rounded to 1024)
    uint8_t* sB = smem + 23552;         // 64 x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 23552 + 8192);
    uint64_t* mma_bar = bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // B = identity, K-major SW128: element (n, k) at n*128 + (((k>>3) ^ (n&7)) << 4) + (k&7)*2
    for (int i = threadIdx.x; i < 64 * 64; i += 128) {
        int n = i >> 6, k = i & 63;
        bf16 v = __float2bfloat16(n == k ? 1.f : 0.f);
        *reinterpret_cast<bf16*>(sB + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = v;
    }
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mma_bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, HR * HC * 128);
        tma_load_4d(sA, &tmA, 0, 0, 0, 0, bar);
    }
    mbar_wait(bar, 0);
    __syncthreads();
    uint32_t phase = 0;
    for (int variant = 0; variant < 2; ++variant) {
        for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, s = tap % 3;
            if (threadIdx.x == 0) {
                tc_fence_after();
                constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
                const uint32_t a_addr = smem_u32(sA) + (r * HC + s) * 128;
                uint64_t da = make_smem_desc(a_addr, 16, HC * 128, 2);
                if (variant == 1) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
                const uint64_t db = make_desc_kmajor_sw128(smem_u32(sB));
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k != 0);
                umma_commit(mma_bar);
            }
            mbar_wait(mma_bar, phase);
            phase ^= 1;
            tc_fence_after();
            const int m = warp * 32 + lane;
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
                tmem_ld_wait();
                for (int j = 0; j < 16; ++j) out[(((size_t)variant * 9 + tap) * 128 + m) * 64 + c0 + j] = __uint_as_float(v[j]);
            }
            tc_fence_before();
            __syncthreads();
        }
    }
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

int main() {
    std::vector<bf16> hx(HR * HC * C);
    std::vector<float> fx(HR * HC * C);
    for (size_t i = 0; i < hx.size(); ++i) { float v = (float)((i * 2654435761u >> 8) % 2001) / 1000.f - 1.f; hx[i] = __float2bfloat16(v); fx[i] = __bfloat162float(hx[i]); }
    bf16* dx; float* dout;
    cudaMalloc(&dx, hx.size() * 2);
    cudaMalloc(&dout, 2 * 9 * 128 * 64 * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    CUtensorMap m;
    uint64_t dims[4] = {C, HC, HR, 1}; uint64_t str[3] = {C * 2, HC * C * 2, (uint64_t)HR * HC * C * 2};
    uint32_t box[4] = {64, HC, HR, 1}; uint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    cudaFuncSetAttribute(halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    halo_kernel<<<1, 128, 40000>>>(m, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> ho(2 * 9 * 128 * 64);
    cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
    for (int variant = 0; variant < 2; ++variant)
        for (int tap = 0; tap < 9; ++tap) {
            int bad = 0, rr = tap / 3, ss = tap % 3;
            for (int mm = 0; mm < 128; ++mm)
                for (int n = 0; n < 64; ++n) {
                    float want = fx[(((mm / 8) + rr) * HC + (mm % 8) + ss) * C + n];
                    float got = ho[(((size_t)variant * 9 + tap) * 128 + mm) * 64 + n];
                    if (want != got) ++bad;
                }
            printf("variant %d (base_offset %s) tap (%d,%d): mismatches %d / 8192\n", variant, variant ? "set" : "0", rr, ss, bad);
        }
    return 0;
}
