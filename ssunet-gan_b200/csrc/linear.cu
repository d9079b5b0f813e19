// Skinny linear layers of the discriminator head (batch rows << features): weight-streaming kernels.
// The weight (fp32 master, [nout][k]) is read once per 8-row batch chunk; activations are tiny and L2-resident.
#include "common.cuh"

namespace ssg {

constexpr int MCH = 8;

// Weight-streaming forward: a block owns JB output features and walks the whole reduction dimension in 4-element
// vectors (float4 weights, 4 x bf16 / float4 activations), MROWS batch rows at a time, so every weight element is read
// once per MROWS rows and feeds MROWS FMAs; block-level tree reduction at the end.  (The first version gave each warp one
// feature and issued one 2-byte activation load per FMA: 0.58 ms for fc1 at batch 16; the weight read alone is 12 us.)
constexpr int LIN_JB = 4, LIN_MROWS = 16, LIN_THREADS = 256;

__device__ __forceinline__ void load4f(const float* p, float* f) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
__device__ __forceinline__ void load4f(const bf16* p, float* f) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
}

template <typename T>
__global__ void __launch_bounds__(LIN_THREADS) linear_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, T* __restrict__ y, int m, int k,
                                                                  int nout, int act, float slope, const float* __restrict__ inv_scale) {
    __shared__ float red[LIN_THREADS / 32][LIN_JB * LIN_MROWS];
    const int j0 = blockIdx.x * LIN_JB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float sc = inv_scale ? inv_scale[0] : 1.f;
    const bool vec = (k % 4 == 0);
    for (int m0 = 0; m0 < m; m0 += LIN_MROWS) {
        float acc[LIN_JB][LIN_MROWS];
#pragma unroll
        for (int j = 0; j < LIN_JB; ++j)
#pragma unroll
            for (int i = 0; i < LIN_MROWS; ++i) acc[j][i] = 0.f;
        if (vec) {
            for (int kk = threadIdx.x * 4; kk < k; kk += LIN_THREADS * 4) {
                float wv[LIN_JB][4];
#pragma unroll
                for (int j = 0; j < LIN_JB; ++j) {
                    if (j0 + j < nout) load4f(w + (long long)(j0 + j) * k + kk, wv[j]);
                    else wv[j][0] = wv[j][1] = wv[j][2] = wv[j][3] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < LIN_MROWS; ++i) {
                    if (m0 + i < m) {
                        float xv[4];
                        load4f(x + (long long)(m0 + i) * k + kk, xv);
#pragma unroll
                        for (int j = 0; j < LIN_JB; ++j)
                            acc[j][i] = fmaf(wv[j][0], xv[0], fmaf(wv[j][1], xv[1], fmaf(wv[j][2], xv[2], fmaf(wv[j][3], xv[3], acc[j][i]))));
                    }
                }
            }
        } else {
            for (int kk = threadIdx.x; kk < k; kk += LIN_THREADS) {
#pragma unroll
                for (int j = 0; j < LIN_JB; ++j) {
                    const float wv = j0 + j < nout ? w[(long long)(j0 + j) * k + kk] : 0.f;
#pragma unroll
                    for (int i = 0; i < LIN_MROWS; ++i)
                        if (m0 + i < m) acc[j][i] = fmaf(wv, to_f(x[(long long)(m0 + i) * k + kk]), acc[j][i]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < LIN_JB; ++j)
#pragma unroll
            for (int i = 0; i < LIN_MROWS; ++i) {
                const float v = warp_sum(acc[j][i]);
                if (lane == 0) red[warp][j * LIN_MROWS + i] = v;
            }
        __syncthreads();
        if (threadIdx.x < LIN_JB * LIN_MROWS) {
            float v = 0.f;
#pragma unroll
            for (int wi = 0; wi < LIN_THREADS / 32; ++wi) v += red[wi][threadIdx.x];
            const int j = j0 + threadIdx.x / LIN_MROWS, i = m0 + threadIdx.x % LIN_MROWS;
            if (j < nout && i < m) {
                v = v * sc + (bias ? bias[j] : 0.f);
                y[(long long)i * nout + j] = from_f<T>(apply_act(v, act, slope));
            }
        }
        __syncthreads();
    }
}

// thread per input feature kk; j range split over blockIdx.y, combined with fp32 atomics into dx32
template <typename T>
__global__ void __launch_bounds__(256) linear_dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w,
                                                            float* __restrict__ dx32, int m, int k, int nout, int j_per_block,
                                                            const float* __restrict__ inv_scale) {
    const int kk = blockIdx.x * blockDim.x + threadIdx.x;
    const int j0 = blockIdx.y * j_per_block;
    const int j1 = min(nout, j0 + j_per_block);
    const float sc = inv_scale ? inv_scale[0] : 1.f;
    if (kk >= k) return;
    for (int m0 = 0; m0 < m; m0 += MCH) {
        float acc[MCH];
#pragma unroll
        for (int i = 0; i < MCH; ++i) acc[i] = 0.f;
        for (int j = j0; j < j1; ++j) {
            const float wv = w[(long long)j * k + kk];
#pragma unroll
            for (int i = 0; i < MCH; ++i)
                if (m0 + i < m) acc[i] = fmaf(wv, to_f(dy[(long long)(m0 + i) * nout + j]), acc[i]);
        }
#pragma unroll
        for (int i = 0; i < MCH; ++i)
            if (m0 + i < m) atomicAdd(&dx32[(long long)(m0 + i) * k + kk], acc[i] * sc);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) linear_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw,
                                                            float* __restrict__ dbias, int m, int k, int nout) {
    const int kk = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (kk >= k) return;
    float acc = 0.f, accb = 0.f;
    for (int i = 0; i < m; ++i) {
        const float d = to_f(dy[(long long)i * nout + j]);
        acc = fmaf(d, to_f(x[(long long)i * k + kk]), acc);
        accb += d;
    }
    dw[(long long)j * k + kk] = acc;
    if (dbias && kk == 0) dbias[j] = accb;
}

}  // namespace ssg
using namespace ssg;

extern "C" {

int ssg_linear_fwd(const void* x, const float* w, const float* bias, void* y, int dtype, int m, int k, int nout, int act,
                   float slope, const float* inv_scale_dev, ssg_stream_t s) {
    SSG_CHECK_ARG(m > 0 && k > 0 && nout > 0, "linear_fwd: bad shape");
    unsigned g = (unsigned)((nout + LIN_JB - 1) / LIN_JB);
    SSG_DISPATCH_DTYPE(dtype, linear_fwd_kernel<T><<<g, LIN_THREADS, 0, (cudaStream_t)s>>>((const T*)x, w, bias, (T*)y, m, k, nout, act, slope, inv_scale_dev));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_linear_dgrad(const void* dy, const float* w, float* dx, int dtype, int m, int k, int nout, const float* inv_scale_dev,
                     ssg_stream_t s) {
    SSG_CHECK_ARG(m > 0 && k > 0 && nout > 0, "linear_dgrad: bad shape");
    SSG_CHECK_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)m * k, (cudaStream_t)s));
    int kb = (k + 255) / 256;
    int splits = (2 * sm_count_cached() + kb - 1) / kb;
    if (splits > nout) splits = nout;
    if (splits < 1) splits = 1;
    int jpb = (nout + splits - 1) / splits;
    dim3 grid((unsigned)kb, (unsigned)((nout + jpb - 1) / jpb));
    SSG_DISPATCH_DTYPE(dtype, linear_dgrad_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)dy, w, (float*)dx, m, k, nout, jpb, inv_scale_dev));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_linear_wgrad(const void* x, const void* dy, float* dw, float* dbias, int dtype, int m, int k, int nout, ssg_stream_t s) {
    SSG_CHECK_ARG(m > 0 && k > 0 && nout > 0 && nout <= 65535, "linear_wgrad: bad shape");
    dim3 grid((unsigned)((k + 255) / 256), (unsigned)nout);
    SSG_DISPATCH_DTYPE(dtype, linear_wgrad_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, (const T*)dy, dw, dbias, m, k, nout));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
