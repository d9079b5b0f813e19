// Skinny linear layers of the discriminator head (batch rows << features): weight-streaming kernels.
// The weight (fp32 master, [nout][k]) is read once per 8-row batch chunk; activations are tiny and L2-resident.
#include "common.cuh"

namespace ssg {

constexpr int MCH = 8;

// one warp per output feature j
template <typename T>
__global__ void __launch_bounds__(256) linear_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, T* __restrict__ y, int m, int k, int nout,
                                                          int act, float slope, const float* __restrict__ inv_scale) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nout) return;
    const float sc = inv_scale ? inv_scale[0] : 1.f;
    const float* wr = w + (long long)warp * k;
    for (int m0 = 0; m0 < m; m0 += MCH) {
        float acc[MCH];
#pragma unroll
        for (int i = 0; i < MCH; ++i) acc[i] = 0.f;
        for (int kk = lane; kk < k; kk += 32) {
            const float wv = wr[kk];
#pragma unroll
            for (int i = 0; i < MCH; ++i)
                if (m0 + i < m) acc[i] = fmaf(wv, to_f(x[(long long)(m0 + i) * k + kk]), acc[i]);
        }
#pragma unroll
        for (int i = 0; i < MCH; ++i) {
            float v = warp_sum(acc[i]);
            if (lane == 0 && m0 + i < m) {
                v = v * sc + (bias ? bias[warp] : 0.f);
                y[(long long)(m0 + i) * nout + warp] = from_f<T>(apply_act(v, act, slope));
            }
        }
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
    unsigned g = (unsigned)(((long long)nout * 32 + 255) / 256);
    SSG_DISPATCH_DTYPE(dtype, linear_fwd_kernel<T><<<g, 256, 0, (cudaStream_t)s>>>((const T*)x, w, bias, (T*)y, m, k, nout, act, slope, inv_scale_dev));
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
