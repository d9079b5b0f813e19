// Experiment: UMMA shared-memory descriptors WITHOUT swizzle, for operands assembled from 16-byte-pitch pixel boxes
// (im2col of 8-channel tensors: one {8 ch, 8 w, 16 h} TMA box per filter tap = 128 pixels x 16 B = 2048 B).
// The smem image is  byte(g, p, c) = g * 2048 + p * 16 + c * 2   (g = tap / channel group, p = pixel, c = channel in group).
//  case MN: A[k = p][m = g * 8 + c], B[k = p][n = g * 8 + c]  (both MN-major, K = 128 pixels, 8 MMAs)   -- weight gradient
//  case K : A[m = p][k = t * 8 + c] (K-major, K = 80 = 10 taps), B[n][k] at (k / 8) * 1024 + n * 16 + (k % 8) * 2  -- forward
// Each case is run with both assignments of (LBO, SBO) to (K-direction, MN-direction) core-matrix strides.
#include "../ssunet-gan_b200/csrc/tc_common.cuh"
#include <stdlib.h>
#include <vector>
using namespace ssg::tc;
typedef __nv_bfloat16 bf16;

__global__ void __launch_bounds__(128) nosw_kernel(const bf16* gA_mn, const bf16* gB_mn, const bf16* gA_k, const bf16* gB_k, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                  // 32 KB
    uint8_t* sB = smem + 32768;          // 16 KB
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + 49152);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(mma_bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    uint32_t phase = 0;
    for (int cs = 0; cs < 2; ++cs) {
        // ---- fill smem for this case ----
        __syncthreads();
        if (cs == 0) {
            for (int i = threadIdx.x; i < 128 * 128; i += 128) {      // A_mn: logical [p][m]
                const int p = i / 128, m = i % 128;
                *reinterpret_cast<bf16*>(sA + (m / 8) * 2048 + p * 16 + (m % 8) * 2) = gA_mn[i];
            }
            for (int i = threadIdx.x; i < 128 * 64; i += 128) {       // B_mn: logical [p][n]
                const int p = i / 64, n = i % 64;
                *reinterpret_cast<bf16*>(sB + (n / 8) * 2048 + p * 16 + (n % 8) * 2) = gB_mn[i];
            }
        } else {
            for (int i = threadIdx.x; i < 128 * 80; i += 128) {       // A_k: logical [m = p][k]
                const int p = i / 80, k = i % 80;
                *reinterpret_cast<bf16*>(sA + (k / 8) * 2048 + p * 16 + (k % 8) * 2) = gA_k[i];
            }
            for (int i = threadIdx.x; i < 64 * 80; i += 128) {        // B_k: logical [n][k]
                const int n = i / 80, k = i % 80;
                *reinterpret_cast<bf16*>(sB + (k / 8) * 1024 + n * 16 + (k % 8) * 2) = gB_k[i];
            }
        }
        fence_proxy_async();
        __syncthreads();
        for (int variant = 0; variant < 2; ++variant) {
            if (threadIdx.x == 0) {
                tc_fence_after();
                if (cs == 0) {
                    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
                    // variant 0: LBO = K-direction (next 8 pixels, 128 B), SBO = MN-direction (next group of 8, 2048 B)
                    const uint32_t lbo = variant == 0 ? 128 : 2048, sbo = variant == 0 ? 2048 : 128;
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t da = make_smem_desc(smem_u32(sA) + k * 256, lbo, sbo, 0);
                        const uint64_t db = make_smem_desc(smem_u32(sB) + k * 256, lbo, sbo, 0);
                        umma_bf16(tmem_base, da, db, idesc, k != 0);
                    }
                } else {
                    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
                    // variant 0: LBO = K-direction (next 8 channels: A 2048 B, B 1024 B), SBO = MN-direction (next 8 rows, 128 B)
                    for (int k = 0; k < 5; ++k) {
                        const uint64_t da = variant == 0 ? make_smem_desc(smem_u32(sA) + k * 4096, 2048, 128, 0)
                                                         : make_smem_desc(smem_u32(sA) + k * 4096, 128, 2048, 0);
                        const uint64_t db = variant == 0 ? make_smem_desc(smem_u32(sB) + k * 2048, 1024, 128, 0)
                                                         : make_smem_desc(smem_u32(sB) + k * 2048, 128, 1024, 0);
                        umma_bf16(tmem_base, da, db, idesc, k != 0);
                    }
                }
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
                for (int j = 0; j < 16; ++j) out[(((size_t)cs * 2 + variant) * 128 + m) * 64 + c0 + j] = __uint_as_float(v[j]);
            }
            tc_fence_before();
            __syncthreads();
        }
    }
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

static float rnd(size_t i) { return (float)((i * 2654435761u >> 8) % 17) / 8.f - 1.f; }   // small exact values

int main() {
    std::vector<bf16> a_mn(128 * 128), b_mn(128 * 64), a_k(128 * 80), b_k(64 * 80);
    std::vector<float> fa_mn(a_mn.size()), fb_mn(b_mn.size()), fa_k(a_k.size()), fb_k(b_k.size());
    auto fill = [](std::vector<bf16>& h, std::vector<float>& f, size_t salt) {
        for (size_t i = 0; i < h.size(); ++i) { h[i] = __float2bfloat16(rnd(i * 7 + salt)); f[i] = __bfloat162float(h[i]); }
    };
    fill(a_mn, fa_mn, 1); fill(b_mn, fb_mn, 2); fill(a_k, fa_k, 3); fill(b_k, fb_k, 4);
    bf16 *d0, *d1, *d2, *d3; float* dout;
    cudaMalloc(&d0, a_mn.size() * 2); cudaMalloc(&d1, b_mn.size() * 2); cudaMalloc(&d2, a_k.size() * 2); cudaMalloc(&d3, b_k.size() * 2);
    cudaMalloc(&dout, 4 * 128 * 64 * 4);
    cudaMemcpy(d0, a_mn.data(), a_mn.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(d1, b_mn.data(), b_mn.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(d2, a_k.data(), a_k.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(d3, b_k.data(), b_k.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(nosw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52000);
    nosw_kernel<<<1, 128, 52000>>>(d0, d1, d2, d3, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> ho(4 * 128 * 64);
    cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
    for (int cs = 0; cs < 2; ++cs)
        for (int variant = 0; variant < 2; ++variant) {
            int bad = 0; double maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    double want = 0;
                    if (cs == 0) { for (int p = 0; p < 128; ++p) want += (double)fa_mn[p * 128 + m] * fb_mn[p * 64 + n]; }
                    else { for (int k = 0; k < 80; ++k) want += (double)fa_k[m * 80 + k] * fb_k[n * 80 + k]; }
                    double got = ho[(((size_t)cs * 2 + variant) * 128 + m) * 64 + n];
                    double err = fabs(got - want);
                    if (err > 1e-3) ++bad;
                    if (err > maxerr) maxerr = err;
                }
            printf("case %s variant %d (%s): mismatches %d / 8192, max err %.4f\n", cs == 0 ? "MN-major" : "K-major ", variant,
                   variant == 0 ? "LBO = K-dir, SBO = MN-dir" : "LBO = MN-dir, SBO = K-dir", bad, maxerr);
        }
    return 0;
}
