// Shared helpers for the ssunet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/ssunet_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "ssunet_b200 kernels are written for sm_100a only"
#endif

namespace ssg {

void set_error(const char* fmt, ...);

#define SSG_CHECK_ARG(cond, ...)                         \
    do {                                                 \
        if (!(cond)) {                                   \
            ::ssg::set_error(__VA_ARGS__);               \
            return SSG_ERR_INVALID_ARG;                  \
        }                                                \
    } while (0)

#define SSG_CHECK_CUDA(expr)                                                         \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            ::ssg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return SSG_ERR_CUDA;                                                     \
        }                                                                            \
    } while (0)

#define SSG_CHECK_LAUNCH() SSG_CHECK_CUDA(cudaGetLastError())

// Dispatch on storage dtype: body sees `T` (float or __nv_bfloat16).
#define SSG_DISPATCH_DTYPE(dt, ...)                                   \
    do {                                                              \
        if ((dt) == SSG_F32) { using T = float; __VA_ARGS__; }        \
        else if ((dt) == SSG_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
        else { ::ssg::set_error("unknown dtype %d", (int)(dt)); return SSG_ERR_INVALID_ARG; } \
    } while (0)

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector of T: 4 floats or 8 bf16.
template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    float4 raw;
    __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
    __device__ __forceinline__ void get(float* f) const { f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w; }
    __device__ __forceinline__ void set(const float* f) { raw = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec<bf16> {
    static constexpr int N = 8;
    uint4 raw;
    __device__ __forceinline__ void load(const bf16* p) { raw = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
    __device__ __forceinline__ void get(float* f) const {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ void set(const float* f) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        raw = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == SSG_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == SSG_ACT_LEAKY) return v > 0.f ? v : v * slope;
    return v;
}
// derivative of the activation expressed through its OUTPUT y (sign-preserving activations)
__device__ __forceinline__ float act_grad_from_out(float y, int act, float slope) {
    if (act == SSG_ACT_RELU) return y > 0.f ? 1.f : 0.f;
    if (act == SSG_ACT_LEAKY) return y > 0.f ? 1.f : slope;
    return 1.f;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

inline int sm_count_cached() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

inline unsigned grid_for(long long work_items, int per_block, int max_waves = 8) {
    long long b = (work_items + per_block - 1) / per_block;
    long long cap = (long long)sm_count_cached() * max_waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace ssg
